// numpy's legacy random stream on the device.  pyswarm draws its random numbers from numpy's global MT19937
// (np.random.rand / np.random.uniform; the reference reaches it at nmrfit/utils.py:176-182), and "the same seeded host
// RNG stream" is what end-to-end parity means - so the parity mode of nmrfit_b200.fit used to draw those numbers on the
// host and ship them: two thirds of a 4 ms fit.  This kernel continues the SAME stream on the GPU: given the generator's
// 624-word state and position it produces the next n doubles exactly as RandomState.random_sample does
// (genrand_res53: (a >> 5) * 2^26 + (b >> 6), over 2^53, from two tempered 32-bit words) and hands back the advanced
// state, which the host puts back with np.random.set_state - the stream ends where pyswarm would have left it.
//
// The state recurrence is sequential from block to block (624 words) but parallel inside a block: word k of the next
// block needs words k, k + 1, k + 397 of the current one, so a block regenerates in four dependent sweeps
// ([0, 227), [227, 454), [454, 623), {623}).  One CTA walks the blocks; a C1 fit (100 particles x 22 parameters x
// 2 x 101 draws) is 1,425 blocks, ~0.1 ms.
#include <cuda_runtime.h>
#include <cstdint>
#include "nmrfit_internal.h"

namespace nmrfit {

namespace {

constexpr int kMtN = 624, kMtM = 397, kMtThreads = 256;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t next, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (next & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// key_io [624] (global): the state, updated in place; *pos_io: words of it already consumed (0..624).
// Doubles q = 0..n-1 of the stream go to out_a[q] - or, when nsd > 0, de-interleaved as pyswarm consumes them
// (uniform(size=(S, D)) for rp, then again for rg, per generation): q = (2 g + which) nsd + e -> (which ? out_b : out_a)[g nsd + e].
__global__ void __launch_bounds__(kMtThreads)
mt19937_kernel(uint32_t* __restrict__ key_io, int* __restrict__ pos_io, long long n, double* __restrict__ out_a,
               double* __restrict__ out_b, long long nsd) {
    __shared__ uint32_t key[kMtN];
    __shared__ uint32_t word[kMtN];                        // tempered words of the current block
    const int tid = threadIdx.x;
    for (int k = tid; k < kMtN; k += kMtThreads) key[k] = key_io[k];
    const int pos = *pos_io;
    __syncthreads();
    const long long n_words = 2 * n;
    // stream word t = 0, 1, ... is the t-th word drawn from now on; t0 = stream index of word 0 of the current block
    long long t0 = pos < kMtN ? -(long long)pos : 0;
    bool twist = pos >= kMtN;                              // the given block is used up: start with a fresh one
    uint32_t carry = 0;                                    // last word of the previous block (a pair may straddle two)
    long long last_t0 = t0;
    while (t0 < n_words) {
        if (twist) {
            // regenerate: four sweeps, each reading only words the previous sweeps have finished with
            for (int k = tid; k < kMtN - kMtM; k += kMtThreads) key[k] = mt_twist(key[k], key[k + 1], key[k + kMtM]);
            __syncthreads();
            for (int k = kMtN - kMtM + tid; k < 2 * (kMtN - kMtM); k += kMtThreads)
                key[k] = mt_twist(key[k], key[k + 1], key[k - (kMtN - kMtM)]);
            __syncthreads();
            for (int k = 2 * (kMtN - kMtM) + tid; k < kMtN - 1; k += kMtThreads)
                key[k] = mt_twist(key[k], key[k + 1], key[k - (kMtN - kMtM)]);
            __syncthreads();
            if (tid == 0) key[kMtN - 1] = mt_twist(key[kMtN - 1], key[0], key[kMtM - 1]);
            __syncthreads();
        }
        twist = true;
        for (int k = tid; k < kMtN; k += kMtThreads) word[k] = mt_temper(key[k]);
        __syncthreads();
        // the doubles whose SECOND word (stream word 2q + 1) lies in this block; the first is here too or is `carry`
        const long long lo = t0 < 0 ? 0 : t0;                                         // first stream word in use here
        const long long hi = (t0 + kMtN < n_words ? t0 + kMtN : n_words) - 1;         // last one
        for (long long q = lo / 2 + tid; 2 * q + 1 <= hi; q += kMtThreads) {
            const long long ta = 2 * q, tb = 2 * q + 1;
            const uint32_t a = ta >= t0 ? word[ta - t0] : carry;
            const uint32_t b = word[tb - t0];
            const double x = ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
            if (nsd > 0) {
                const long long blk = q / nsd, e = q - blk * nsd;
                ((blk & 1) ? out_b : out_a)[(blk >> 1) * nsd + e] = x;
            } else {
                out_a[q] = x;
            }
        }
        __syncthreads();
        carry = word[kMtN - 1];
        last_t0 = t0;
        t0 += kMtN;
        __syncthreads();
    }
    for (int k = tid; k < kMtN; k += kMtThreads) key_io[k] = key[k];
    if (tid == 0 && n_words > 0) *pos_io = (int)(n_words - last_t0);      // words consumed from the last block touched
}

}  // namespace

cudaError_t launch_mt19937(unsigned* key_dev, int* pos_dev, long long n, double* out_a, double* out_b, long long nsd,
                           cudaStream_t st) {
    mt19937_kernel<<<1, kMtThreads, 0, st>>>(key_dev, pos_dev, n, out_a, out_b, nsd);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace nmrfit
