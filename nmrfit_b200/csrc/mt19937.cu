// numpy's legacy random stream on the device.  pyswarm draws its random numbers from numpy's global MT19937
// (np.random.rand / np.random.uniform; the reference reaches it at nmrfit/utils.py:176-182), and "the same seeded host
// RNG stream" is what end-to-end parity means - so the parity mode of nmrfit_b200.fit used to draw those numbers on the
// host and ship them: two thirds of a 4 ms fit.  This kernel continues the SAME stream on the GPU: given the generator's
// 624-word state and position it produces the next n doubles exactly as RandomState.random_sample does
// (genrand_res53: (a >> 5) * 2^26 + (b >> 6), over 2^53, from two tempered 32-bit words) and hands back the advanced
// state, which the host puts back with np.random.set_state - the stream ends where pyswarm would have left it.
//
// The recurrence x[n] = f(x[n-624], x[n-623], x[n-227]) is sequential with a lag of 227 words: one CTA advances 227
// words per step (one __syncthreads each) and writes the tempered words out; a second, fully parallel kernel turns word
// pairs into doubles.  A C1 fit (100 particles x 22 parameters x 2 x 101 draws = 888,800 words) is ~3,900 steps.
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

// Pass 1 (sequential in steps, one CTA): the raw recurrence x[n] = twist(x[n-624], x[n-623], x[n-227]) advances
// 227 words per step - the lag of the nearest dependence - with ONE __syncthreads per step; every new word is tempered
// on the spot and written to the word stream.  key_io [624] (global): the state, updated in place; *pos_io: words of it
// already consumed (0..624); words_out [n_words]: the next n_words tempered 32-bit words of the stream.
constexpr int kMtLag = kMtN - kMtM;                        // 227
__global__ void __launch_bounds__(kMtThreads)
mt19937_words_kernel(uint32_t* __restrict__ key_io, int* __restrict__ pos_io, long long n_words,
                     uint32_t* __restrict__ words_out) {
    __shared__ uint32_t ring[1024];                        // x[n] at slot n & 1023 (needs the last 624 + 227 words)
    const int tid = threadIdx.x;
    const int pos = *pos_io;
    // absolute numbering: the given state is x[0..623]; the stream's word t (t = 0, 1, ...) is temper(x[pos + t])
    for (int k = tid; k < kMtN; k += kMtThreads) ring[k] = key_io[k];
    __syncthreads();
    const long long last = (long long)pos + n_words;       // one past the last word needed (absolute)
    for (long long n = pos + tid; n < (last < kMtN ? last : kMtN); n += kMtThreads)
        words_out[n - pos] = mt_temper(ring[n]);           // what is left of the given block
    long long base = kMtN;
    while (base < last) {
        const long long n = base + tid;
        if (tid < kMtLag && n < last + kMtN) {              // (a little past the end, so that a whole final block exists)
            const uint32_t x = mt_twist(ring[(n - kMtN) & 1023], ring[(n - kMtN + 1) & 1023], ring[(n - kMtLag) & 1023]);
            ring[n & 1023] = x;
            if (n < last) words_out[n - pos] = mt_temper(x);
        }
        base += kMtLag;
        __syncthreads();
    }
    // the state to hand back: numpy keeps whole blocks, so the block that holds the last word drawn, [b0, b0 + 624).
    // If that block was only partly generated above, finish it (its words are never drawn here, only stored).
    const long long b0 = last <= kMtN ? 0 : ((last - 1) / kMtN) * kMtN;
    long long done = base;                                 // words [0, done) exist... up to last + 624 at most
    while (done < b0 + kMtN) {
        const long long n = done + tid;
        if (tid < kMtLag && n < b0 + kMtN)
            ring[n & 1023] = mt_twist(ring[(n - kMtN) & 1023], ring[(n - kMtN + 1) & 1023], ring[(n - kMtLag) & 1023]);
        done += kMtLag;
        __syncthreads();
    }
    for (int k = tid; k < kMtN; k += kMtThreads) key_io[k] = ring[(b0 + k) & 1023];
    if (tid == 0 && n_words > 0) *pos_io = (int)(last - b0);
}

// Pass 2 (fully parallel): doubles from word pairs - genrand_res53 - de-interleaved as pyswarm consumes them.
// Double q uses words (2q, 2q + 1); with nsd > 0, q = (2 g + which) nsd + e -> (which ? out_b : out_a)[g nsd + e].
__global__ void mt19937_doubles_kernel(const uint32_t* __restrict__ words, long long n, double* __restrict__ out_a,
                                       double* __restrict__ out_b, long long nsd) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t a = words[2 * q], b = words[2 * q + 1];
    const double x = ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
    if (nsd > 0) {
        const long long blk = q / nsd, e = q - blk * nsd;
        ((blk & 1) ? out_b : out_a)[(blk >> 1) * nsd + e] = x;
    } else {
        out_a[q] = x;
    }
}

}  // namespace

cudaError_t launch_mt19937(unsigned* key_dev, int* pos_dev, long long n, unsigned* words_dev, double* out_a, double* out_b,
                           long long nsd, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    mt19937_words_kernel<<<1, kMtThreads, 0, st>>>(key_dev, pos_dev, 2 * n, words_dev);
    mt19937_doubles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(words_dev, n, out_a, out_b, nsd);
    count_launches(2);
    return cudaGetLastError();
}

}  // namespace nmrfit
