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

constexpr int kMtN = 624, kMtM = 397, kMtThreads = 128;

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
// 227 words per step - the lag of the nearest dependence - with ONE __syncthreads per step; every new word goes to the
// word stream as it is (the second pass tempers).  key_io [624] (global): the state, updated in place; *pos_io: words of it
// already consumed (0..624); words_out [n_words]: the next n_words 32-bit words of the stream, UNTEMPERED.
constexpr int kMtLag = kMtN - kMtM;                        // 227
// (128 threads, two words each: the step is a chain of shared-memory round trips and a barrier - four warps keep the
// barrier cheap and give every thread two independent words to overlap the latencies; 32-bit ring arithmetic.)
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
    // The UNTEMPERED words go out (the parallel second pass tempers them): every instruction saved here is saved ~4,000
    // times in a row on one warp scheduler - the kernel is a single chain of dependent steps.
    for (long long n = pos + tid; n < (last < kMtN ? last : kMtN); n += kMtThreads)
        words_out[n - pos] = ring[n];                      // what is left of the given block
    // the state to hand back: numpy keeps whole blocks, so the block that holds the last word drawn, [b0, b0 + 624);
    // the recurrence runs to its end (words past `last` are stored in the ring, not written out)
    const long long b0 = last <= kMtN ? 0 : ((last - 1) / kMtN) * kMtN;
    const long long end = b0 + kMtN;
    const bool second = tid + kMtThreads < kMtLag;         // this thread's second word of a step exists
    // steps [0, full): every word of the step exists and is written out; the last few steps are checked word by word
    const long long n_steps = (end - kMtN + kMtLag - 1) / kMtLag;
    const long long full_ll = last > kMtN ? (last - kMtN) / kMtLag : 0;
    const int steps = (int)n_steps, full = (int)(full_ll < n_steps ? full_ll : n_steps);
    unsigned r0 = (unsigned)tid;                           // (n - 624) & 1023 of this thread's first word: n starts at 624 + tid
    uint32_t* o0 = words_out + (kMtN - pos) + tid;         // where the first word of the current step goes
    for (int st = 0; st < full; ++st) {
        const unsigned r1 = (r0 + kMtThreads) & 1023u;
        const uint32_t x0 = mt_twist(ring[r0], ring[(r0 + 1) & 1023u], ring[(r0 + kMtM) & 1023u]);
        uint32_t x1 = 0;
        if (second) x1 = mt_twist(ring[r1], ring[(r1 + 1) & 1023u], ring[(r1 + kMtM) & 1023u]);
        // (slots read in a step - 624 + 227 consecutive ones - and slots written never coincide modulo 1,024)
        ring[(r0 + kMtN) & 1023u] = x0;
        o0[0] = x0;
        if (second) {
            ring[(r1 + kMtN) & 1023u] = x1;
            o0[kMtThreads] = x1;
        }
        r0 = (r0 + kMtLag) & 1023u;
        o0 += kMtLag;
        __syncthreads();
    }
    for (int st = full; st < steps; ++st) {
        const long long n0 = kMtN + (long long)st * kMtLag + tid, n1 = n0 + kMtThreads;
        const bool do0 = n0 < end, do1 = second && n1 < end;
        const unsigned r1 = (r0 + kMtThreads) & 1023u;
        uint32_t x0 = 0, x1 = 0;
        if (do0) x0 = mt_twist(ring[r0], ring[(r0 + 1) & 1023u], ring[(r0 + kMtM) & 1023u]);
        if (do1) x1 = mt_twist(ring[r1], ring[(r1 + 1) & 1023u], ring[(r1 + kMtM) & 1023u]);
        if (do0) ring[(r0 + kMtN) & 1023u] = x0;
        if (do1) ring[(r1 + kMtN) & 1023u] = x1;
        if (do0 && n0 < last) o0[0] = x0;
        if (do1 && n1 < last) o0[kMtThreads] = x1;
        r0 = (r0 + kMtLag) & 1023u;
        o0 += kMtLag;
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
    const uint32_t a = mt_temper(words[2 * q]), b = mt_temper(words[2 * q + 1]);     // (pass 1 writes raw words)
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
