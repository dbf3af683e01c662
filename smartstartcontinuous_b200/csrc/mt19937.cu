// mt19937.cu -- numpy's legacy RandomState stream (MT19937), generated on the device.
//
// The reference draws the K*H*da action samples of a decision with
//     npr.uniform(self.low, self.high, (self.N, self.horizon, da))          NND_MB_agent.py:500-501
// i.e. low + (high - low) * random_sample() from the global Mersenne Twister, element by element.
// At K = 131072, H = 50 that is 33 ms of host time and a 52 MB upload for a 2.4 ms decision.  Here
// the SAME stream is produced on the GPU, bit for bit: the host hands over the 624-word key and the
// position, the kernel returns the samples in device memory and the key / position after the draw.
//
// MT19937 is a sequential recurrence (x[k+624] = x[k+397] ^ twist(x[k], x[k+1])), parallel only 227
// words at a time.  Parallelism across SMs comes from jump-ahead (Haramoto, Matsumoto, Nishimura,
// Panneton, L'Ecuyer 2008): the state transition is linear over GF(2) with a characteristic polynomial
// phi of degree 19937, so the window J steps ahead is
//     window_J[j] = XOR_{i : g_i = 1} x[i + j],        g = x^J mod phi,
// a GF(2) correlation of the next 19937 + 624 raw words with the bits of g.  The stream is cut into
// P segments of whole 624-word blocks; a segment's jump is evaluated by one CTA or -- medium streams, few
// segments, spare SMs -- by several CTAs that each take a slice of g and combine through global memory (g
// depends on the segment start only, so the polynomials are computed once per batch shape on the host and
// cached on the device; the layout is chosen by a cost model, make_layout).  The segment's CTA then regenerates
// its blocks exactly like genrand does and hands the raw words to the TMA engine; a second, fully parallel
// kernel tempers them and converts pairs of words into doubles ((a >> 5) * 2^26 + (b >> 6)) / 2^53
// (randomkit's rk_double) and applies low + (high - low) * u with the reference's two roundings.
//
// phi is held as its 135 exponents (derived by Berlekamp-Massey from numpy's own output:
// oracle/mt19937_poly.py; tests/test_mt19937_poly.py re-derives it and checks the table).
#include <array>
#include <atomic>
#include <cstdlib>

#include "mpc_kernels.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397, MT_DEG = 19937;
constexpr int POLY_W64 = 312;                 // 19968 bits
constexpr int MT_L = MT_N - MT_M;             // 227: the recurrence's lag
constexpr int MT_THREADS = 640;               // 20 warps
constexpr int GEN_WARPS = 8;                  // warps 0..7 run the recurrence, 8..19 temper / convert / store
constexpr int CONV_PER_LANE = 20;             // jump: lane t owns outputs 20 t .. 20 t + 19 (32 x 20 = 640 >= 625)
constexpr int CONV_OUT = 32 * CONV_PER_LANE;
constexpr int CONV_WARPS = MT_THREADS / 32;   // every warp takes the coefficient words w = warp, warp + 20, ...
constexpr int SEQ_BLOCKS = 34;                // raw words a jump reads: i < 19968, j < 640, + 16-byte load slack
constexpr int SEQ_WORDS = SEQ_BLOCKS * MT_N;

// exponents of phi below 19937
const uint16_t PHI_LOW[134] = {
    0, 1189, 1416, 1585, 1643, 1870, 2493, 2773, 3000, 3227, 3454, 3681, 3908, 4135, 4362, 4753, 5661, 6337,
    6569, 7129, 7477, 7525, 7583, 7752, 7979, 8206, 9505, 9901, 9969, 10128, 10693, 10761, 10920, 11089, 11147,
    11157, 11215, 11321, 11374, 11384, 11485, 11611, 11712, 11717, 11838, 11881, 11944, 11997, 12277, 12335,
    12393, 12504, 12509, 12620, 12673, 12731, 12736, 12789, 12905, 12958, 12963, 13137, 13185, 13190, 13243,
    13301, 13412, 13528, 13533, 13639, 13697, 13760, 13813, 13866, 14093, 14151, 14209, 14320, 14325, 14436,
    14547, 14552, 14605, 14721, 14774, 14779, 14953, 15001, 15006, 15059, 15117, 15228, 15344, 15349, 15455,
    15513, 15576, 15629, 15682, 15909, 15967, 16025, 16136, 16141, 16252, 16363, 16368, 16421, 16537, 16590,
    16595, 16817, 16822, 16875, 16933, 17044, 17160, 17271, 17329, 17445, 17498, 17725, 17783, 17841, 17952,
    18068, 18179, 18237, 18406, 18633, 18691, 18860, 19087, 19314};

// ------------------------------------------------------------------ GF(2)[x] mod phi on the host
using Poly = std::array<uint64_t, POLY_W64>;

inline void xor_shifted(uint64_t* a, uint64_t w, int s) {
    const int q = s >> 6, r = s & 63;
    a[q] ^= w << r;
    if (r) a[q + 1] ^= w >> (64 - r);
}

// a: 2 * POLY_W64 words (degree < 39936) -> reduced into the low 19937 bits.  phi's second term is
// x^19314, 623 below the leading one, so folding a whole 64-bit word never lands in the word itself.
void poly_reduce(uint64_t* a) {
    for (int i = 2 * POLY_W64 - 1; i >= POLY_W64; --i) {
        const uint64_t w = a[i];
        if (!w) continue;
        a[i] = 0;
        const int base = 64 * i - MT_DEG;
        for (uint16_t e : PHI_LOW) xor_shifted(a, w, base + e);
    }
    const int top = MT_DEG - 64 * (POLY_W64 - 1);          // bit 33 of the last word is x^19937
    const uint64_t w = a[POLY_W64 - 1] >> top;
    if (w) {
        a[POLY_W64 - 1] &= (1ull << top) - 1;
        for (uint16_t e : PHI_LOW) xor_shifted(a, w, e);
    }
}

Poly poly_sqr(const Poly& a) {
    static uint16_t spread[256];
    static bool init = false;
    if (!init) {
        for (int v = 0; v < 256; ++v) {
            uint16_t s = 0;
            for (int b = 0; b < 8; ++b)
                if (v >> b & 1) s |= (uint16_t)(1u << (2 * b));
            spread[v] = s;
        }
        init = true;
    }
    std::vector<uint64_t> t(2 * POLY_W64 + 1, 0);
    for (int i = 0; i < POLY_W64; ++i) {
        uint64_t lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) {
            lo |= (uint64_t)spread[(a[i] >> (8 * b)) & 255] << (16 * b);
            hi |= (uint64_t)spread[(a[i] >> (8 * b + 32)) & 255] << (16 * b);
        }
        t[2 * i] = lo;
        t[2 * i + 1] = hi;
    }
    poly_reduce(t.data());
    Poly r;
    std::copy(t.begin(), t.begin() + POLY_W64, r.begin());
    return r;
}

Poly poly_mul(const Poly& a, const Poly& b) {
    // 64 shifted copies of b, then one contiguous XOR per set bit of a
    std::vector<uint64_t> sh((size_t)64 * (POLY_W64 + 1), 0);
    for (int s = 0; s < 64; ++s) {
        uint64_t* row = sh.data() + (size_t)s * (POLY_W64 + 1);
        for (int k = 0; k < POLY_W64; ++k) {
            row[k] |= b[k] << s;
            if (s) row[k + 1] |= b[k] >> (64 - s);
        }
    }
    std::vector<uint64_t> t(2 * POLY_W64 + 1, 0);
    for (int i = 0; i < POLY_W64; ++i) {
        uint64_t w = a[i];
        while (w) {
            const int s = __builtin_ctzll(w);
            w &= w - 1;
            const uint64_t* row = sh.data() + (size_t)s * (POLY_W64 + 1);
            uint64_t* dst = t.data() + i;
            for (int k = 0; k <= POLY_W64; ++k) dst[k] ^= row[k];
        }
    }
    poly_reduce(t.data());
    Poly r;
    std::copy(t.begin(), t.begin() + POLY_W64, r.begin());
    return r;
}

// x^J mod phi: left-to-right square and multiply by x
Poly poly_pow_x(uint64_t J) {
    Poly g{};
    g[0] = 1;
    if (J == 0) return g;
    const int top_word = POLY_W64 - 1, top_bit = MT_DEG - 64 * (POLY_W64 - 1);
    for (int bit = 63 - __builtin_clzll(J); bit >= 0; --bit) {
        g = poly_sqr(g);
        if (J >> bit & 1) {
            uint64_t carry = 0;
            for (int i = 0; i < POLY_W64; ++i) {
                const uint64_t n = g[i] >> 63;
                g[i] = (g[i] << 1) | carry;
                carry = n;
            }
            if (g[top_word] >> top_bit & 1) {
                g[top_word] &= (1ull << top_bit) - 1;
                for (uint16_t e : PHI_LOW) g[e >> 6] ^= 1ull << (e & 63);
            }
        }
    }
    return g;
}

// ------------------------------------------------------------------ device side
struct MtKey { uint32_t w[MT_N]; };

struct MtSegment {
    long long first_block;     // index of the first 624-word block (block 0 = the key itself)
    long long end_block;       // one past the last block whose doubles this segment emits
    int poly;                  // row of the polynomial table (jump to 624 * first_block - 1), -1: no jump
    int emit;                  // 0: state-only segment
    int write_state;           // this segment may write the final state
    // a jump shared by `share_n` CTAs: CTA `share_r` correlates the coefficient words m with m % share_n ==
    // share_r (in units of 20 words), XORs its 625 partial outputs into window `window` in global memory and
    // counts itself in; share_r == 0 waits for everybody and goes on to generate the segment, the others exit
    int share_r, share_n, window;
};

struct MtArgs {
    const MtSegment* segs;
    const uint32_t* polys;     // [n_polys][624] coefficient words of g (bit i of word w = x^(32 w + i))
    uint32_t* raw;             // untempered words, block by block from block `raw_first_block`
    long long raw_first_block;
    long long total_w;         // stream words of the whole draw (all shards)
    int pos;                   // numpy's pos: words of the key block already consumed (0..624)
    uint32_t* state_out;       // mapped host memory: [0..1] flag (u64), [2] pos, [4 .. 4 + 624) key
    unsigned long long seq;
    uint32_t* windows;         // [n_windows][640] partial jump results (zero between calls) ...
    unsigned int* counters;    // ... and how many CTAs have contributed to each
};

// raw words -> doubles: element e of the shard = words (w, w + 1), w = word_base + 2 e, of the raw buffer
struct MtConvertArgs {
    const uint32_t* raw;
    double* out;
    long long word_base;
    long long count;
    long long first_elem;      // global element index of out[0] (for the period of low / range)
    int period;
    double low[SS_MAX_DA], range[SS_MAX_DA];
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((b & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// One block of genrand's in-place update, written out of place and WITHOUT a barrier inside: thread
// t < 227 owns n[t], n[t + 227] and n[t + 454]; each is the previous one of the same thread XOR a twist
// of old words, and the single cross-thread input (n[0] for n[623]) is recomputed by its consumer.
__device__ __forceinline__ uint32_t mt_next0(const uint32_t* c) { return c[MT_M] ^ mt_twist(c[0], c[1]); }
__device__ __forceinline__ void mt_next_block(const uint32_t* __restrict__ c, uint32_t* __restrict__ n, int t) {
    if (t >= MT_L) return;
    // every load up front (one shared-memory round trip), then three dependent XORs
    const bool third = t < MT_N - 2 * MT_L;
    const int k = third ? t + 2 * MT_L : 0;
    const uint32_t a0 = c[t], a1 = c[t + 1], am = c[t + MT_M];
    const uint32_t b0 = c[t + MT_L], b1 = c[t + MT_L + 1];
    const uint32_t c0 = c[k], c1 = c[k + 1 < MT_N ? k + 1 : 0];
    const uint32_t z0 = c[0], z1 = c[1], zm = c[MT_M];
    const uint32_t v0 = am ^ mt_twist(a0, a1);
    const uint32_t v1 = v0 ^ mt_twist(b0, b1);
    n[t] = v0;
    n[t + MT_L] = v1;
    if (third) {
        const uint32_t nxt = k + 1 < MT_N ? c1 : (zm ^ mt_twist(z0, z1));
        n[k] = v1 ^ mt_twist(c0, nxt);
    }
}

// One pair of coefficient bits (warp-uniform) against the lane's 20 outputs: o[r] ^= g_lo s[r] ^ g_hi s[r + 1].
// Three straight-line variants per bit position keep the unrolled code of a whole coefficient word at
// ~18 KB, inside the 32 KB L1.5 instruction cache (a 16-way nibble switch per position is 52 KB and
// starves the issue slots on instruction fetch: measured).
template <int BASE>
__device__ __forceinline__ void conv_dibit(uint32_t d, uint32_t (&o)[CONV_PER_LANE], const uint32_t (&s)[52]) {
    if (d == 0u) return;
    if (d == 1u) {
#pragma unroll
        for (int r = 0; r < CONV_PER_LANE; ++r) o[r] ^= s[BASE + r];
    } else if (d == 2u) {
#pragma unroll
        for (int r = 0; r < CONV_PER_LANE; ++r) o[r] ^= s[BASE + 1 + r];
    } else {
#pragma unroll
        for (int r = 0; r < CONV_PER_LANE; ++r) o[r] ^= s[BASE + r] ^ s[BASE + 1 + r];
    }
}

template <int POS>
__device__ __forceinline__ void conv_word(uint32_t gw, uint32_t (&o)[CONV_PER_LANE], const uint32_t (&s)[52]) {
    if constexpr (POS < 16) {
        conv_dibit<2 * POS>((gw >> (2 * POS)) & 3u, o, s);
        conv_word<POS + 1>(gw, o, s);
    }
}

// shared memory: [seq 34 x 624 | conv partials 20 x 640 | g 624 | block ring 8 x 624]
constexpr int RING_SLOTS = 8;
constexpr int SMEM_WORDS = SEQ_WORDS + CONV_WARPS * CONV_OUT + MT_N + RING_SLOTS * MT_N;

__global__ void __launch_bounds__(MT_THREADS, 1)
mt19937_raw_kernel(const __grid_constant__ MtKey key, const __grid_constant__ MtArgs args) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* seq = smem;
    uint32_t* part = seq + SEQ_WORDS;
    uint32_t* g_s = part + CONV_WARPS * CONV_OUT;
    uint32_t* ring = g_s + MT_N;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const MtSegment seg = args.segs[blockIdx.x];
    pdl_trigger();       // the conversion kernel may be scheduled (it waits for this grid to complete)

    if (seg.poly < 0) {
        // the segment starts at the key block itself
        for (int i = tid; i < MT_N; i += MT_THREADS) ring[i] = key.w[i];
        __syncthreads();
    } else {
        // ---- raw words x[0 .. SEQ_WORDS) continuing the key, one 624-word block per barrier
        for (int i = tid; i < MT_N; i += MT_THREADS) seq[i] = key.w[i];
        const uint32_t* gp = args.polys + (size_t)seg.poly * MT_N;
        for (int i = tid; i < MT_N; i += MT_THREADS) g_s[i] = gp[i];
        __syncthreads();
        for (int blk = 0; blk + 1 < SEQ_BLOCKS; ++blk) {
            mt_next_block(seq + blk * MT_N, seq + (blk + 1) * MT_N, tid);
            __syncthreads();
        }
        // ---- correlation with g.  Warp = a set of coefficient words, lane t = outputs 20 t .. 20 t + 19:
        // the 51 raw words one coefficient word touches come in with 13 aligned 16-byte loads, and every
        // pair of bits of g (warp-uniform) selects one of three straight-line XOR patterns.
        {
            uint32_t o[CONV_PER_LANE];
#pragma unroll
            for (int r = 0; r < CONV_PER_LANE; ++r) o[r] = 0u;
            const uint32_t* sp = seq + CONV_PER_LANE * lane;
            for (int w = warp + CONV_WARPS * seg.share_r; w < MT_N; w += CONV_WARPS * seg.share_n) {
                const uint32_t gw = g_s[w];
                if (gw == 0u) continue;
                uint32_t s[52];
                const uint4* q = reinterpret_cast<const uint4*>(sp + 32 * w);
#pragma unroll
                for (int v = 0; v < 13; ++v) {
                    const uint4 x = q[v];
                    s[4 * v] = x.x; s[4 * v + 1] = x.y; s[4 * v + 2] = x.z; s[4 * v + 3] = x.w;
                }
                conv_word<0>(gw, o, s);
            }
            uint4* dst = reinterpret_cast<uint4*>(part + warp * CONV_OUT + CONV_PER_LANE * lane);
#pragma unroll
            for (int v = 0; v < CONV_PER_LANE / 4; ++v) dst[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        }
        __syncthreads();
        // conv word j = x[J + j], J = 624 * first_block - 1: the block is words 1 .. 624
        uint32_t v = 0;
        if (tid < MT_N) {
            const int j = tid + 1;
#pragma unroll
            for (int p = 0; p < CONV_WARPS; ++p) v ^= part[p * CONV_OUT + j];
        }
        if (seg.share_n > 1) {
            // several CTAs share this jump: combine through global memory (XOR is order-free), count in, and
            // let share 0 carry on once everybody has contributed (all CTAs of the grid are co-resident:
            // cooperative launch)
            uint32_t* win = args.windows + (size_t)seg.window * CONV_OUT;
            unsigned int* cnt = args.counters + seg.window;
            if (tid < MT_N) atomicXor(win + tid, v);
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(cnt, 1u);
            if (seg.share_r != 0) return;
            if (tid == 0) {
                int spins = 0;
                while (*reinterpret_cast<volatile unsigned int*>(cnt) < (unsigned int)seg.share_n)
                    if (++spins > (1 << 26)) asm volatile("trap;");
                __threadfence();
            }
            __syncthreads();
            if (tid < MT_N) {
                v = __ldcg(win + tid);
                win[tid] = 0u;                       // zero again for the next call
            }
            if (tid == 0) *cnt = 0u;
        }
        if (tid < MT_N) ring[tid] = v;
        __syncthreads();
    }

    // ---- block after block (warps 0..7; the others are done).  The untempered words leave through
    // the TMA engine (one 2496-byte bulk store per block, issued by one thread): per-thread global
    // stores in front of the per-block barrier would put their L2 round trip on the recurrence's
    // critical path (measured: 0.25 instead of 0.19 us per block).
    if (warp >= GEN_WARPS) return;
    const long long last_w = (long long)args.pos + args.total_w - 1;      // last stream word of the draw
    const long long state_block = last_w / MT_N;
    const int nb = (int)(seg.end_block - seg.first_block);
    const int state_it = seg.write_state && args.state_out && state_block >= seg.first_block &&
                                 state_block < seg.end_block ? (int)(state_block - seg.first_block) : -1;
    uint32_t* gdst = args.raw + (seg.first_block - args.raw_first_block) * MT_N;
    const bool store = seg.emit != 0;
    // ring slot 0 holds the first block (written with generic stores above, barrier since)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(32 * GEN_WARPS) : "memory");
    int cur = 0;
    for (int it = 0; it < nb; ++it) {
        const uint32_t* c_blk = ring + cur * MT_N;
        if (store && tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                         "r"((uint32_t)__cvta_generic_to_shared(c_blk)), "n"(MT_N * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        gdst += MT_N;
        if (it == state_it) {
            for (int i = tid; i < MT_N; i += 32 * GEN_WARPS) args.state_out[4 + i] = c_blk[i];
            if (tid == 0) args.state_out[2] = (uint32_t)(last_w % MT_N) + 1u;
            __threadfence_system();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * GEN_WARPS) : "memory");
            if (tid == 0) {
                *reinterpret_cast<volatile unsigned long long*>(args.state_out) = args.seq;
                __threadfence_system();
            }
        }
        if (it + 1 < nb) {
            const int nxt = cur + 1 == RING_SLOTS ? 0 : cur + 1;
            mt_next_block(c_blk, ring + nxt * MT_N, tid);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            cur = nxt;
        }
        // block it + 2 will overwrite the slot of block it + 2 - RING_SLOTS: of the it + 1 bulk stores
        // committed so far all but the newest RING_SLOTS - 2 must have finished reading shared memory
        if (store && tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(RING_SLOTS - 2) : "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * GEN_WARPS) : "memory");
    }
    if (store && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// temper, randomkit's rk_double ((a >> 5) * 2^26 + (b >> 6)) / 2^53, then low + (high - low) * u with
// the two roundings of the reference's float64 expression.  The integer -> double conversions are
// exact bit constructions: 2^25 + a 2^-27 and 1/2 + b 2^-53 have a and b as their low mantissa words.
__global__ void __launch_bounds__(256)
mt19937_convert_kernel(const __grid_constant__ MtConvertArgs a) {
    pdl_wait();          // programmatic dependent of the generator kernel
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < a.count; e += stride) {
        const uint32_t* w = a.raw + a.word_base + 2 * e;
        const uint32_t hi = mt_temper(w[0]) >> 5, lo = mt_temper(w[1]) >> 6;
        const double r = (__hiloint2double(0x41800000, (int)hi) - 33554432.0) + (__hiloint2double(0x3FE00000, (int)lo) - 0.5);
        int j = 0;
        if (a.period > 1) j = (int)((a.first_elem + e) % a.period);
        a.out[e] = __dadd_rn(__dmul_rn(r, a.range[j]), a.low[j]);
    }
}

// ------------------------------------------------------------------ per-shape plans (cached)
struct MtPlan {
    long long off_w = -1, n_w = 0, total_w = 0;
    long long b_lo = 0, n_blocks = 0;
    int n_segments = 0, n_windows = 0;
    bool cooperative = false;      // jumps shared between CTAs: the grid must be co-resident
    DevBuf segs, polys;
};

struct MtCache {
    std::vector<MtPlan> plans;
    DevBuf raw, windows;           // windows: [n][640] partial jump results + [n] counters, zero between calls
    size_t windows_n = 0;
    void* host_state = nullptr;
    void* host_state_dev = nullptr;
    unsigned long long seq = 0;
    bool attr_set = false;
};

MtCache* cache_of(ss_ctx* c) {
    if (!c->mt_cache) c->mt_cache = new MtCache();
    return static_cast<MtCache*>(c->mt_cache);
}

// one way to cut the blocks [b_lo, b_hi] into P segments: CTA descriptors, polynomials and the modelled time
struct MtLayout {
    std::vector<MtSegment> segs;       // one entry per CTA
    std::vector<uint32_t> words;
    int n_windows = 0;
    bool shared_jumps = false;
    double cost_us = 0.0;
};

// measured on B200: 0.19 us per regenerated block, 5.7 us for the 34 raw blocks a jump reads, 56 us for the
// correlation with a dense polynomial (~10 000 set bits; zero coefficients are skipped, and x^J mod phi stays
// sparse for J up to a few times 19937 because phi has only 135 terms), ~1.5 us for the hand-off of a shared jump
double layout_cost(long long per, int n_jumps, int shares, int max_weight) {
    if (n_jumps == 0) return 0.19 * (double)per;
    return 0.19 * (double)per + 5.7 + 56.0 * max_weight / 10000.0 / shares + (shares > 1 ? 1.5 : 0.0);
}

// with_polys = false: shape and cost only (dense polynomials assumed)
MtLayout make_layout(long long b_lo, long long b_hi, long long total_w, int P, int sm_count, bool with_polys) {
    MtLayout L;
    const long long n_blocks = b_hi - b_lo + 1;
    const long long per = (n_blocks + P - 1) / P;
    std::vector<MtSegment> logical;
    std::vector<uint64_t> jumps;
    for (long long b = b_lo; b <= b_hi; b += per) {
        MtSegment s{};
        s.first_block = b;
        s.end_block = std::min(b + per, b_hi + 1);
        s.emit = 1;
        s.write_state = 1;
        s.poly = -1;
        if (b > 0) { s.poly = (int)jumps.size(); jumps.push_back((uint64_t)b * MT_N - 1); }
        logical.push_back(s);
    }
    // the block that holds the state after the WHOLE draw (all shards): one of two, depending on pos
    const long long f0 = (total_w - 1) / MT_N;
    if (!(b_lo <= f0 && f0 + 1 <= b_hi)) {
        for (auto& s : logical) s.write_state = 0;
        MtSegment s{};
        s.first_block = f0;
        s.end_block = f0 + 2;
        s.emit = 0;
        s.write_state = 1;
        s.poly = -1;
        if (f0 > 0) { s.poly = (int)jumps.size(); jumps.push_back((uint64_t)f0 * MT_N - 1); }
        logical.push_back(s);
    }
    // polynomials: the regular segments are `per` blocks apart -> one power, then a chain of products
    int max_weight = jumps.empty() ? 0 : 10000;
    if (with_polys) {
        L.words.resize(jumps.size() * (size_t)MT_N);
        Poly step{}, g{};
        bool have_step = false;
        max_weight = 0;
        for (size_t i = 0; i < jumps.size(); ++i) {
            const bool chain = i > 0 && jumps[i] - jumps[i - 1] == (uint64_t)per * MT_N;
            if (chain) {
                if (!have_step) { step = poly_pow_x((uint64_t)per * MT_N); have_step = true; }
                g = poly_mul(g, step);
            } else {
                g = poly_pow_x(jumps[i]);
            }
            int weight = 0;
            for (int k = 0; k < POLY_W64; ++k) {
                L.words[i * MT_N + 2 * k] = (uint32_t)g[k];
                L.words[i * MT_N + 2 * k + 1] = (uint32_t)(g[k] >> 32);
                weight += __builtin_popcountll(g[k]);
            }
            max_weight = std::max(max_weight, weight);
        }
    }
    // spare SMs share the jumps: every jump gets `shares` CTAs, each correlating a part of the polynomial
    const int n_jumps = (int)jumps.size(), n_plain = (int)logical.size() - n_jumps;
    int shares = 1;
    if (n_jumps > 0) shares = std::max(1, std::min(MT_N / CONV_WARPS, (sm_count - n_plain) / n_jumps));
    if (n_jumps > 0 && 56.0 * max_weight / 10000.0 < 4.0) shares = 1;          // not worth a hand-off
    for (const MtSegment& lg : logical) {
        const int n = lg.poly >= 0 ? shares : 1;
        for (int r = 0; r < n; ++r) {
            MtSegment s = lg;
            s.share_r = r;
            s.share_n = n;
            s.window = lg.poly >= 0 ? lg.poly : 0;
            L.segs.push_back(s);
        }
    }
    L.n_windows = n_jumps;
    L.shared_jumps = shares > 1;
    L.cost_us = layout_cost(per, n_jumps, shares, max_weight);
    return L;
}

int build_plan(ss_ctx* c, MtPlan& p, long long off_w, long long n_w, long long total_w) {
    p.off_w = off_w; p.n_w = n_w; p.total_w = total_w;
    const long long b_lo = off_w / MT_N;
    const long long b_hi = (MT_N + off_w + n_w - 1) / MT_N;        // last block a position of 624 can reach
    const long long n_blocks = b_hi - b_lo + 1;
    p.b_lo = b_lo;
    p.n_blocks = n_blocks;
    MtLayout best;
    if (std::getenv("SS_MT_FORCE_P")) {
        best = make_layout(b_lo, b_hi, total_w, std::max(1, std::atoi(std::getenv("SS_MT_FORCE_P"))), c->sm_count, true);
    } else {
        // candidates: few segments (cheap sparse jumps, each shared by many SMs) ... one segment per SM (dense
        // jumps, nothing to share).  Small P are scored with their actual polynomial weights, the large ones
        // as dense; the winner's polynomials are computed last.
        const int cand[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 74, 148};
        int best_P = 1;
        bool have = false, best_has_polys = false;
        for (int P : cand) {
            if (P > c->sm_count || (P > 1 && n_blocks / P < 4)) break;
            if (n_blocks >= 4096 && P < 48) continue;                      // long streams: dense either way
            const bool exact = P <= 32;
            MtLayout L = make_layout(b_lo, b_hi, total_w, P, c->sm_count, exact);
            if (!have || L.cost_us < best.cost_us) {
                best = std::move(L);
                best_P = P;
                best_has_polys = exact;
                have = true;
            }
        }
        if (!best_has_polys) best = make_layout(b_lo, b_hi, total_w, best_P, c->sm_count, true);
    }
    p.n_segments = (int)best.segs.size();
    p.cooperative = best.shared_jumps;
    p.n_windows = best.n_windows;
    SS_CUDA_CHECK(c, p.segs.ensure(best.segs.size() * sizeof(MtSegment)));
    SS_CUDA_CHECK(c, cudaMemcpyAsync(p.segs.p, best.segs.data(), best.segs.size() * sizeof(MtSegment), cudaMemcpyHostToDevice, c->stream));
    if (!best.words.empty()) {
        SS_CUDA_CHECK(c, p.polys.ensure(best.words.size() * 4));
        SS_CUDA_CHECK(c, cudaMemcpyAsync(p.polys.p, best.words.data(), best.words.size() * 4, cudaMemcpyHostToDevice, c->stream));
    }
    SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));       // the layout is a local
    return SS_OK;
}

}  // namespace

void mt19937_release(ss_ctx* c) {
    if (!c->mt_cache) return;
    MtCache* m = static_cast<MtCache*>(c->mt_cache);
    for (auto& p : m->plans) { p.segs.release(); p.polys.release(); }
    m->raw.release();
    m->windows.release();
    if (m->host_state) cudaFreeHost(m->host_state);
    delete m;
    c->mt_cache = nullptr;
}

extern "C" int ss_mt19937_jump_poly(uint64_t J, uint32_t* out_words) {
    if (!out_words) return SS_EINVAL;
    const Poly g = poly_pow_x(J);
    for (int k = 0; k < POLY_W64; ++k) {
        out_words[2 * k] = (uint32_t)g[k];
        out_words[2 * k + 1] = (uint32_t)(g[k] >> 32);
    }
    return SS_OK;
}

extern "C" int ss_mt19937_phi_exponents(int* out, int cap) {
    const int n = (int)(sizeof(PHI_LOW) / sizeof(PHI_LOW[0])) + 1;
    if (out)
        for (int i = 0; i < n && i < cap; ++i) out[i] = i + 1 < n ? PHI_LOW[i] : MT_DEG;
    return n;
}

extern "C" int ss_mt19937_uniform(ss_ctx* c, const uint32_t* key, int pos, int64_t n_total, int64_t first,
                                  int64_t count, int period, const double* low, const double* high,
                                  double** out_dev) {
    if (!c) return SS_EINVAL;
    if (!key || pos < 0 || pos > MT_N || n_total < 1 || first < 0 || count < 1 || first + count > n_total ||
        period < 1 || period > SS_MAX_DA || !low || !high || !out_dev)
        SS_FAIL(c, SS_EINVAL, "mt19937: bad arguments (key, 0 <= pos <= 624, 0 <= first, first + count <= n_total, 1 <= period <= 8)");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    MtCache* m = cache_of(c);
    if (!m->host_state) {
        SS_CUDA_CHECK(c, cudaHostAlloc(&m->host_state, 4096, cudaHostAllocMapped));
        std::memset(m->host_state, 0, 4096);
        SS_CUDA_CHECK(c, cudaHostGetDevicePointer(&m->host_state_dev, m->host_state, 0));
    }
    if (!m->attr_set) {
        SS_CUDA_CHECK(c, cudaFuncSetAttribute(mt19937_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SMEM_WORDS * 4));
        m->attr_set = true;
    }
    const long long off_w = 2 * first, n_w = 2 * count, total_w = 2 * n_total;
    MtPlan* plan = nullptr;
    for (auto& p : m->plans)
        if (p.off_w == off_w && p.n_w == n_w && p.total_w == total_w) plan = &p;
    if (!plan) {
        if (m->plans.size() >= 16) {
            m->plans.front().segs.release();
            m->plans.front().polys.release();
            m->plans.erase(m->plans.begin());
        }
        m->plans.emplace_back();
        plan = &m->plans.back();
        int rc = build_plan(c, *plan, off_w, n_w, total_w);
        if (rc) { m->plans.pop_back(); return rc; }
    }
    SS_CUDA_CHECK(c, c->mpc_actions64.ensure((size_t)count * 8));
    SS_CUDA_CHECK(c, m->raw.ensure((size_t)plan->n_blocks * MT_N * 4 + 16));
    MtKey k;
    std::memcpy(k.w, key, sizeof(k.w));
    MtArgs a{};
    a.segs = plan->segs.as<MtSegment>();
    a.polys = plan->polys.as<uint32_t>();
    a.raw = m->raw.as<uint32_t>();
    a.raw_first_block = plan->b_lo;
    a.total_w = total_w;
    a.pos = pos;
    a.state_out = reinterpret_cast<uint32_t*>(m->host_state_dev);
    a.seq = ++m->seq;
    if ((size_t)plan->n_windows > m->windows_n) {
        const size_t n = (size_t)plan->n_windows;
        SS_CUDA_CHECK(c, m->windows.ensure(n * (CONV_OUT + 1) * 4));
        SS_CUDA_CHECK(c, cudaMemsetAsync(m->windows.p, 0, n * (CONV_OUT + 1) * 4, c->stream));
        m->windows_n = n;
    }
    a.windows = m->windows.as<uint32_t>();
    a.counters = reinterpret_cast<unsigned int*>(m->windows.as<uint32_t>() + m->windows_n * CONV_OUT);
    {
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(plan->n_segments);
        cfg.blockDim = dim3(MT_THREADS);
        cfg.dynamicSmemBytes = SMEM_WORDS * 4;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = plan->cooperative ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SS_CUDA_CHECK(c, cudaLaunchKernelEx(&cfg, mt19937_raw_kernel, k, a));
    }
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    MtConvertArgs cv{};
    cv.raw = a.raw;
    cv.out = c->mpc_actions64.as<double>();
    cv.word_base = pos + off_w - plan->b_lo * MT_N;
    cv.count = count;
    cv.first_elem = first;
    cv.period = period;
    for (int j = 0; j < SS_MAX_DA; ++j) {
        cv.low[j] = j < period ? low[j] : 0.0;
        cv.range[j] = j < period ? high[j] - low[j] : 0.0;
    }
    const long long want_blocks = (count + 255) / 256;
    const int grid = (int)std::min<long long>(want_blocks, (long long)c->sm_count * 8);
    SS_CUDA_CHECK(c, launch_dependent(mt19937_convert_kernel, dim3(grid), dim3(256), 0, c->stream, cv));
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    *out_dev = cv.out;
    return SS_OK;
}

extern "C" int ss_mt19937_state(ss_ctx* c, uint32_t* out_key, int* out_pos) {
    if (!c) return SS_EINVAL;
    if (!c->mt_cache || !static_cast<MtCache*>(c->mt_cache)->seq) SS_FAIL(c, SS_ESTATE, "mt19937: no draw to take the state of");
    if (!out_key || !out_pos) SS_FAIL(c, SS_EINVAL, "mt19937: null output");
    MtCache* m = static_cast<MtCache*>(c->mt_cache);
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(m->host_state);
    unsigned spins = 0;
    while (*flag != m->seq) {
        if ((++spins & 0x3fff) == 0) {
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) SS_CUDA_CHECK(c, e);
            if (e == cudaSuccess && *flag != m->seq) {
                SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
                if (*flag != m->seq) SS_FAIL(c, SS_ECUDA, "mt19937: the kernel finished without writing the state");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const uint32_t* h = reinterpret_cast<const uint32_t*>(m->host_state);
    *out_pos = (int)h[2];
    std::memcpy(out_key, h + 4, MT_N * 4);
    return SS_OK;
}

// get_best_sim_actions with the reference's own draw in ONE call: the samples of npr.uniform(low, high,
// (K_global, H, da)) (NND_MB_agent.py:500-501) generated on the device from numpy's generator state, the decision
// on them (ss_mpc_plan), and the generator state after the draw written back through mt_key / mt_pos -- which may
// point straight at numpy's own state struct.
extern "C" int ss_mpc_plan_mt19937(ss_ctx* c, const double* state, int wp_index, int64_t K_local, int64_t k_offset,
                                   int64_t K_global, int H, uint32_t* mt_key, int* mt_pos, const double* act_low,
                                   const double* act_high, double gamma, double hpf, int penalty_mode, int precision,
                                   int64_t* out_best_k, double* out_best_score, double* out_best_sequence,
                                   double* out_best_path, double* out_scores) {
    if (!c) return SS_EINVAL;
    if (!mt_key || !mt_pos) SS_FAIL(c, SS_EINVAL, "mt19937: null generator state");
    if (!c->model_set) SS_FAIL(c, SS_ESTATE, "mpc: model and plan must be set before planning");
    const int64_t per_seq = (int64_t)H * c->da;
    double* dev = nullptr;
    int rc = ss_mt19937_uniform(c, mt_key, *mt_pos, K_global * per_seq, k_offset * per_seq, K_local * per_seq, c->da,
                                act_low, act_high, &dev);
    if (rc) return rc;
    rc = ss_mpc_plan(c, state, wp_index, K_local, k_offset, K_global, H, dev, 0, act_low, act_high, gamma, hpf,
                     penalty_mode, precision, out_best_k, out_best_score, out_best_sequence, out_best_path, out_scores);
    if (rc) return rc;
    return ss_mt19937_state(c, mt_key, mt_pos);
}
