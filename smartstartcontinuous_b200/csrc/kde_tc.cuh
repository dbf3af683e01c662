// kde_tc.cuh -- tcgen05 variant of the KDE pair kernel (included by kde.cu; d <= 20).
//
// The CUDA-core pair kernel is dispatch-bound: every kernel evaluation costs ~6 issue cycles for
// the exponent (differences or the expanded form) before the MUFU.EX2 it is nominally limited by.
// Here the exponent  e'[q][i] = -|q|^2 - |x_i|^2 + 2 q.x_i  of a 128-query x 256-point block comes
// out of KS / 16 tcgen05.mma instructions (M = 128, N = 256, K = KS = 32 / 64 / 128, BF16 in, FP32 accumulate
// in TMEM): every FP32 operand is split into three BF16 pieces (hi, mid, lo; 24 bits) and the six
// significant partial products per dimension get their own K slot, so the product is FP32-accurate
// although it runs on the tensor pipe:
//
//   K slots, dimension j (6j .. 6j+5):  q side  [qh, qh, qm, qh, ql, qm]
//                                       x side  [xh, xm, xh, xl, xh, xm]      (x = 2y)
//   slots 6D .. 6D+2:                   q side  split3(-|q|^2),  x side  1, 1, 1
//   slots 6D+3 .. 6D+5:                 q side  1, 1, 1,         x side  split3(-|x|^2)
//
// The 16 epilogue warps then only have to read e' from TMEM (tcgen05.ld, lane = query), take
// exp2 and add: 2 issue cycles per evaluation on the MUFU path.  That leaves enough issue slots to
// send KDE_TC_POLY of every 8 evaluations through a degree-4 polynomial exp2 on the FMA pipe
// (exp2_poly2, kde.cu), which lifts the kernel ABOVE the MUFU.EX2 ceiling of 16 evaluations / clk
// / SM.  Precision: as the expanded CUDA-core variant (the cancellation in the FP32 accumulator is
// the same), guarded by the same max|y|^2 limit; the caller falls back to the difference kernel
// when the guard trips.
#pragma once
// (kde.cu includes <cuda_bf16.h> at global scope before this header)

namespace kdetc {

constexpr int QT = 128;             // queries per work item (MMA M, TMEM lanes)
constexpr int NP = 256;             // points per tile (MMA N, FP32 columns per accumulator buffer)
// K slots (BF16) of the MMA: 6 d + 6 must fit -> KS = 32 (d <= 4), 64 (d <= 9), 128 (d <= 20)
__host__ __device__ constexpr int ks_for(int d) { return 6 * d + 6 <= 32 ? 32 : (6 * d + 6 <= 64 ? 64 : 128); }
__host__ __device__ constexpr int max_d_for(int ks) { return (ks - 6) / 6; }
__host__ __device__ constexpr int a_bytes(int ks) { return QT * ks * 2; }     // 8 / 16 / 32 KB per query tile
__host__ __device__ constexpr int b_bytes(int ks) { return NP * ks * 2; }     // 16 / 32 / 64 KB per point tile
__host__ __device__ constexpr int nstage_for(int ks) { return ks <= 64 ? 4 : 2; }
constexpr int EPI_WARPS = 16;       // warp w: TMEM lane quarter w % 4, column group w / 4
constexpr int CG_COLS = NP / (EPI_WARPS / 4);   // 64 columns per warp and tile
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int MAX_D = 20;           // largest state dimension of the tcgen05 variant (KS = 128)
#ifndef SS_KDE_TC_POLY_MASK
#define SS_KDE_TC_POLY_MASK 0x5252  // bits 1, 4, 6 (+8): six of every sixteen evaluations take the polynomial
#endif
constexpr unsigned POLY_MASK = SS_KDE_TC_POLY_MASK;   // over the column index mod 16

__device__ __forceinline__ void tc_commit1(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void umma1_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one1() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
// K-major, no swizzle: 8-row x 16-byte core matrices; image [k/8][rows][8]: LBO = rows * 16 B
// (between the K halves of one MMA), SBO = 128 B (between 8-row groups); descriptor version 1
__device__ __forceinline__ uint64_t make_desc1(uint32_t smem_addr, uint32_t rows) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((rows * 16) >> 4) << 16) |
           ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
// D f32, A/B bf16, K-major, M = 128, N = 256
constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);

__device__ __forceinline__ void tmem_ld32x(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// wait for the outstanding tcgen05.ld's; the registers are threaded through the asm so the compiler
// cannot schedule their consumers above the wait
__device__ __forceinline__ void tmem_wait_ld_regs(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                   "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                   "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                   "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// exp2 of 32 exponents, summed: POLY_MASK picks the evaluations that take the FMA-pipe polynomial
// (two at a time, packed f32x2), the others go through MUFU.EX2
__device__ __forceinline__ void exp_sum32(const uint32_t (&v)[32], float (&acc)[4], float2& accp) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (!((POLY_MASK >> (i & 15)) & 1u)) acc[i & 3] += ex2_approx(__uint_as_float(v[i]));
    float pe[32];
    int np = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if ((POLY_MASK >> (i & 15)) & 1u) pe[np++] = __uint_as_float(v[i]);
#pragma unroll
    for (int i = 0; i + 1 < np; i += 2) accp = __fadd2_rn(accp, exp2_poly2(make_float2(pe[i], pe[i + 1])));
    if (np & 1) acc[0] += ex2_approx(pe[np - 1]);
}

// ---- operand images ----------------------------------------------------------------------
// hi / mid / lo BF16 pieces of an FP32 value (hi + mid + lo == v exactly for normal v)
__device__ __forceinline__ void split3(float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// One thread per row.  POINTS: tiles of NP rows, image [tile][KS/8][NP][8]; queries: tiles of QT rows.
// Rows >= n are padding: points far away (-|x|^2 = -1e30), queries zero.
template <bool POINTS, int KS>
__device__ __forceinline__ void whiten_tc_rows(long long block, const double* __restrict__ x, long long n,
                                               long long n_pad, int d, KdeFit* __restrict__ fit,
                                               __nv_bfloat16* __restrict__ img) {
    constexpr int ROWS = POINTS ? NP : QT;
    constexpr int MAX_D = max_d_for(KS);
    const long long i = block * (long long)blockDim.x + threadIdx.x;
    float norm2 = 0.f;
    if (i < n_pad) {
        __nv_bfloat16 slot[KS];
        const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
#pragma unroll
        for (int s = 0; s < KS; ++s) slot[s] = zero;
        float yv[MAX_D];
#pragma unroll
        for (int j = 0; j < MAX_D; ++j) yv[j] = 0.f;
        if (i < n) {
            double cdiff[MAX_D];
#pragma unroll
            for (int j = 0; j < MAX_D; ++j) cdiff[j] = j < d ? x[i * d + j] - fit->mean[j] : 0.0;
            double nn = 0.0;
#pragma unroll
            for (int j = 0; j < MAX_D; ++j) {
                double y = 0.0;
                if (j < d)
#pragma unroll
                    for (int k = 0; k < MAX_D; ++k)
                        if (k <= j) y += fit->wm[j * d + k] * cdiff[k];
                yv[j] = (float)y;
                nn += (double)yv[j] * (double)yv[j];
            }
            norm2 = (float)nn;
        }
        const int base_n = 6 * d;
#pragma unroll
        for (int j = 0; j < MAX_D; ++j) {
            if (j < d) {
                __nv_bfloat16 h, m, l;
                split3(POINTS ? 2.f * yv[j] : yv[j], h, m, l);
                if (POINTS) {
                    slot[6 * j] = h; slot[6 * j + 1] = m; slot[6 * j + 2] = h;
                    slot[6 * j + 3] = l; slot[6 * j + 4] = h; slot[6 * j + 5] = m;
                } else {
                    slot[6 * j] = h; slot[6 * j + 1] = h; slot[6 * j + 2] = m;
                    slot[6 * j + 3] = h; slot[6 * j + 4] = l; slot[6 * j + 5] = m;
                }
            }
        }
        {
            __nv_bfloat16 h, m, l;
            split3(i < n ? -norm2 : (POINTS ? -1e30f : 0.f), h, m, l);
            // runtime slot positions (d is not a template parameter): select without indexing
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const int r = s - base_n;
                if (POINTS) {
                    if (r >= 0 && r < 3) slot[s] = one;
                    if (r == 3) slot[s] = h;
                    if (r == 4) slot[s] = m;
                    if (r == 5) slot[s] = l;
                } else {
                    if (r == 0) slot[s] = h;
                    if (r == 1) slot[s] = m;
                    if (r == 2) slot[s] = l;
                    if (r >= 3 && r < 6) slot[s] = one;
                }
            }
        }
        const long long tile = i / ROWS;
        const int row = (int)(i - tile * ROWS);
        __nv_bfloat16* base = img + (size_t)tile * (ROWS * KS);
#pragma unroll
        for (int kc = 0; kc < KS / 8; ++kc) {
            uint4 v;
            __nv_bfloat162 p0 = __halves2bfloat162(slot[8 * kc], slot[8 * kc + 1]);
            __nv_bfloat162 p1 = __halves2bfloat162(slot[8 * kc + 2], slot[8 * kc + 3]);
            __nv_bfloat162 p2 = __halves2bfloat162(slot[8 * kc + 4], slot[8 * kc + 5]);
            __nv_bfloat162 p3 = __halves2bfloat162(slot[8 * kc + 6], slot[8 * kc + 7]);
            v.x = *reinterpret_cast<uint32_t*>(&p0);
            v.y = *reinterpret_cast<uint32_t*>(&p1);
            v.z = *reinterpret_cast<uint32_t*>(&p2);
            v.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(base + (size_t)kc * (ROWS * 8) + (size_t)row * 8) = v;
        }
    }
    for (int off = 16; off > 0; off >>= 1) norm2 = fmaxf(norm2, __shfl_xor_sync(0xffffffffu, norm2, off));
    if ((threadIdx.x & 31) == 0 && norm2 > 0.f) atomicMax(&fit->max_norm2_bits, __float_as_int(norm2));
}

// points (blocks [0, point_blocks)) and queries (the remaining blocks) in one launch
template <int KS>
__global__ void __launch_bounds__(256)
kde_whiten_tc_kernel(const double* __restrict__ data, long long n, long long n_pad, __nv_bfloat16* __restrict__ p_img,
                     const double* __restrict__ queries, long long m, long long m_pad,
                     __nv_bfloat16* __restrict__ q_img, int d, KdeFit* __restrict__ fit, unsigned point_blocks) {
    pdl_trigger();       // the pair kernel may set itself up
    pdl_wait();          // the fit comes from the moments kernel in front
    if (blockIdx.x < point_blocks)
        whiten_tc_rows<true, KS>(blockIdx.x, data, n, n_pad, d, fit, p_img);
    else
        whiten_tc_rows<false, KS>(blockIdx.x - point_blocks, queries, m, m_pad, d, fit, q_img);
}

// ---- the pair kernel -----------------------------------------------------------------------
// Persistent CTAs; work item = (query tile, slice of point tiles); slices <= n_tiles.
// partial[slice * 4 + column group][m_pad].
template <int KS>
__global__ void __launch_bounds__(THREADS, 1)
kde_pairs_tc_kernel(const __nv_bfloat16* __restrict__ q_img, const __nv_bfloat16* __restrict__ p_img,
                    long long n_tiles, int q_tiles, int slices, long long m_pad,
                    const KdeFit* __restrict__ fit, float* __restrict__ partial) {
    constexpr int A_BYTES = a_bytes(KS), B_BYTES = b_bytes(KS), NSTAGE = nstage_for(KS);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* b_ring = smem;                                   // NSTAGE point tiles
    unsigned char* a_buf = smem + (size_t)NSTAGE * B_BYTES;         // 2 query tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_buf + 2 * A_BYTES);
    uint64_t* b_full = bars;                 // [NSTAGE] TMA -> MMA
    uint64_t* b_empty = b_full + NSTAGE;     // [NSTAGE] MMA (commit) -> TMA
    uint64_t* acc_full = b_empty + NSTAGE;   // [2] MMA (commit) -> epilogue
    uint64_t* acc_free = acc_full + 2;       // [2] epilogue (16 warps) -> MMA
    uint64_t* a_full = acc_free + 2;         // [2] TMA -> MMA
    uint64_t* a_free = a_full + 2;           // [2] MMA (commit) -> TMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_free + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long items = (long long)q_tiles * slices;
    pdl_trigger();                           // the finish kernel may be scheduled (it waits for this grid)
    // the guard is evaluated by the caller after the call (optimistic launch); nothing to do here
    (void)fit;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1); mbar_init(&acc_free[s], EPI_WARPS);
            mbar_init(&a_full[s], 1); mbar_init(&a_free[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    // launched as a programmatic dependent of the whitening kernel: barriers and TMEM are set up while it still
    // runs; nothing it writes (operand images) and nothing the previous consumers read (partial) is touched before
    pdl_wait();

    if (warp < EPI_WARPS) {
        // ================================ EPILOGUE =======================================
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CG_COLS);
        uint32_t g = 0;                     // g: tiles consumed so far (ring / buffer parity)
        for (long long item = blockIdx.x; item < items; item += gridDim.x) {
            const int qt = (int)(item % q_tiles);
            const long long sl = item / q_tiles;
            const long long t0 = sl * n_tiles / slices, t1 = (sl + 1) * n_tiles / slices;   // balanced slices
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            float2 accp = make_float2(0.f, 0.f);
            for (long long t = t0; t < t1; ++t, ++g) {
                const uint32_t b = g & 1;
                mbar_wait(&acc_full[b], (g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t v0[32], v1[32];
                tmem_ld32x(lane_addr + b * NP, v0);
                tmem_ld32x(lane_addr + b * NP + 32, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_free[b])) : "memory");
                }
                exp_sum32(v0, acc, accp);
                exp_sum32(v1, acc, accp);
            }
            // every (slice, column group) pair is its own row of `partial` (the finish kernel sums
            // them): the warps never meet at a barrier and drift freely across tiles and items
            const float tot = (acc[0] + acc[1]) + (acc[2] + acc[3]) + (accp.x + accp.y);
            partial[((size_t)sl * (EPI_WARPS / 4) + cg) * m_pad + (size_t)qt * QT + q * 32 + lane] = tot;
        }
    } else if (warp == EPI_WARPS) {
        // ================================ MMA ISSUER ======================================
        uint32_t g = 0, it_local = 0;
        const uint32_t ring_addr = smem_u32(b_ring), a_addr = smem_u32(a_buf);
        for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it_local) {
            const long long sl = item / q_tiles;
            const long long t0 = sl * n_tiles / slices, t1 = (sl + 1) * n_tiles / slices;
            const uint32_t ia = it_local & 1;
            mbar_wait(&a_full[ia], (it_local >> 1) & 1);
            for (long long t = t0; t < t1; ++t, ++g) {
                const uint32_t st = g % NSTAGE, b = g & 1;
                mbar_wait(&acc_free[b], ((g >> 1) & 1) ^ 1);
                mbar_wait(&b_full[st], (g / NSTAGE) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one1()) {
#pragma unroll
                    for (int ks = 0; ks < KS / 16; ++ks)
                        umma1_ss(tmem + b * NP, make_desc1(a_addr + ia * A_BYTES + ks * 2 * (QT * 16), QT),
                                 make_desc1(ring_addr + st * B_BYTES + ks * 2 * (NP * 16), NP), IDESC1, ks);
                    tc_commit1(&b_empty[st]);
                    tc_commit1(&acc_full[b]);
                    if (t == t1 - 1) tc_commit1(&a_free[ia]);
                }
                __syncwarp();
            }
            if (t1 <= t0 && elect_one1()) tc_commit1(&a_free[ia]);      // empty slice (cannot happen, kept safe)
        }
    } else {
        // ================================ TMA PRODUCER ====================================
        if (lane == 0) {
            uint32_t g = 0, it_local = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it_local) {
                const int qt = (int)(item % q_tiles);
                const long long sl = item / q_tiles;
                const long long t0 = sl * n_tiles / slices, t1 = (sl + 1) * n_tiles / slices;
                const uint32_t ia = it_local & 1;
                mbar_wait(&a_free[ia], ((it_local >> 1) & 1) ^ 1);
                mbar_expect_tx(&a_full[ia], A_BYTES);
                bulk_g2s(a_buf + ia * A_BYTES, reinterpret_cast<const unsigned char*>(q_img) + (size_t)qt * A_BYTES,
                         A_BYTES, &a_full[ia]);
                for (long long t = t0; t < t1; ++t, ++g) {
                    const uint32_t st = g % NSTAGE;
                    mbar_wait(&b_empty[st], ((g / NSTAGE) & 1) ^ 1);
                    mbar_expect_tx(&b_full[st], B_BYTES);
                    bulk_g2s(b_ring + (size_t)st * B_BYTES,
                             reinterpret_cast<const unsigned char*>(p_img) + (size_t)t * B_BYTES, B_BYTES, &b_full[st]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

__host__ constexpr size_t smem_bytes(int ks) { return (size_t)nstage_for(ks) * b_bytes(ks) + 2 * a_bytes(ks) + 16 * 8 + 16 + 128; }

}  // namespace kdetc
