// py_random.cu -- CPython's random.sample(range(first, first + n), k) on the host, in C++.
//
// The reference picks its smart-start candidates with the standard library's global generator:
//     np.array(random.sample(range(first_buffer_index, len(self.buffer)), number_of_states))
//                                                                      replay_buffer.py:152
// In CPython that is ~0.6 us per drawn index (8-13 ms for the 16 384 candidates of BASELINE config 2:
// 20 x the whole GPU selection).  This file restates the algorithm -- Random.sample's pool / set
// branches, _randbelow_with_getrandbits, getrandbits(k <= 32) = genrand_uint32() >> (32 - k), MT19937 --
// on the generator state random.getstate() exposes, so the same indices come out in the same order and
// the global generator is left exactly where random.sample would leave it.  replay_buffer.py verifies
// the restatement against the interpreter's own random.sample once per process and keeps using the
// interpreter if they ever disagree.  No GPU work.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/ss_b200.h"

namespace {

constexpr int N = 624, M = 397;

struct PyMt {
    uint32_t* mt;
    int index;
    void refill() {
        int kk;
        uint32_t y;
        for (kk = 0; kk < N - M; ++kk) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
        }
        for (; kk < N - 1; ++kk) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
        }
        y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (-(y & 1u) & 0x9908b0dfu);
        index = 0;
    }
    static uint32_t temper(uint32_t y) {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint32_t next() {
        if (index >= N) refill();
        return temper(mt[index++]);
    }
    // Random._randbelow_with_getrandbits(n), 0 < n < 2^32
    uint32_t below(uint32_t n, int bits) {
        uint32_t r = next() >> (32 - bits);
        while (r >= n) r = next() >> (32 - bits);
        return r;
    }
};

}  // namespace

extern "C" int ss_py_random_sample(uint32_t* mt_state, int* mt_index, int64_t first, int64_t n, int64_t k,
                                   int64_t* out) {
    if (!mt_state || !mt_index || !out || n < 1 || n >= (1ll << 31) || k < 0 || k > n || *mt_index < 0 || *mt_index > N)
        return SS_EINVAL;
    PyMt g{mt_state, *mt_index};
    // setsize = 21; if k > 5: setsize += 4 ** _ceil(_log(k * 3, 4))
    double setsize = 21.0;
    if (k > 5) setsize += std::pow(4.0, std::ceil(std::log((double)(k * 3)) / std::log(4.0)));
    if ((double)n <= setsize) {
        std::vector<int64_t> pool((size_t)n);
        for (int64_t i = 0; i < n; ++i) pool[(size_t)i] = first + i;
        for (int64_t i = 0; i < k; ++i) {
            const uint32_t m = (uint32_t)(n - i);
            int bits = 0;
            while ((m >> bits) != 0) ++bits;                   // m.bit_length()
            const uint32_t j = g.below(m, bits);
            out[i] = pool[j];
            pool[j] = pool[(size_t)(n - i - 1)];
        }
    } else {
        // selected = set(); per index: j = randbelow(n) until j not in selected.  Both rejections (r >= n and
        // "already selected") only skip generator outputs, so the result is the first k distinct values of
        // the stream of outputs below n.  One 624-word block at a time: temper + shift (a loop the compiler
        // vectorises), branch-free compaction of the outputs below n, then the bitmap pass -- 2 x faster than
        // one unpredictable branch per output.
        std::vector<uint64_t> seen((size_t)((n + 63) / 64), 0);
        int bits = 0;
        while (((uint32_t)n >> bits) != 0) ++bits;
        const int shift = 32 - bits;
        const uint32_t un = (uint32_t)n;
        uint32_t shifted[N], cand[N];
        uint16_t at[N];
        int64_t taken = 0;
        while (taken < k) {
            if (g.index >= N) g.refill();
            for (int i = 0; i < N; ++i) shifted[i] = PyMt::temper(g.mt[i]) >> shift;
            int c = 0;
            for (int i = g.index; i < N; ++i) {
                cand[c] = shifted[i];
                at[c] = (uint16_t)i;
                c += shifted[i] < un;
            }
            g.index = N;
            for (int q = 0; q < c; ++q) {
                const uint32_t j = cand[q];
                uint64_t& word = seen[j >> 6];
                const uint64_t bit = 1ull << (j & 63);
                if (word & bit) continue;
                word |= bit;
                out[taken++] = first + j;
                if (taken == k) {
                    g.index = at[q] + 1;       // the generator stops behind the output that completed the sample
                    break;
                }
            }
        }
    }
    *mt_index = g.index;
    return SS_OK;
}
