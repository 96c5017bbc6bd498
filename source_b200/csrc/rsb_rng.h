// rsb_rng.h -- random streams for the trace loop.
//
//  * Mt19937_64: bit-exact restatement of the reference's global generator
//    (raysect/core/math/random.pyx:99-265: init_genrand64, init_by_array64, _rand_uint64, seed,
//    uniform).  The reference has ONE process-global stream; here every pixel owns a stream
//    seeded exactly as `seed(seed_base + pixel_index)` would, so a frame is independent of the
//    pixel -> thread -> GPU mapping and can be reproduced on the reference with a RenderEngine that
//    re-seeds before each pixel task.  State (312 x u64) lives in HBM, word-interleaved across
//    threads (word i of stream s at state[i*stride + s]) so that a warp's accesses coalesce.
//  * Philox4x32-10: counter-based generator keyed on (seed, pixel, sample) for throughput runs;
//    no state in memory.  Same uniform() mapping: 53 random bits * 2^-53.
#pragma once
#include "rsb_math.h"

namespace rsb {

#define RSB_MT_NN 312
#define RSB_MT_MM 156

#define RSB_MT_WIN_DRAWS 6      // draws covered by a preloaded window (a Lambert bounce + the next roulette take 5)
#define RSB_MT_WIN_WORDS 13     // words i .. i+6 and i+156 .. i+161 of the ring
#define RSB_MT_WIN_STRIDE 128   // window element k of a thread at win[k * 128] (shared memory, one column per thread)

struct Mt19937_64 {
    uint64_t* mt;       // base of this stream's words
    size_t stride;      // distance (in words) between consecutive words of the stream
    int mti;
    // Optional preloaded window (device kernels): the state words the next RSB_MT_WIN_DRAWS draws read, fetched
    // from HBM in ONE batch of independent loads instead of one dependent round trip per draw.  The values are
    // exactly what the draws would load themselves: draw k reads words i+k, i+k+1 (both still old) and i+k+156
    // (mod 312; never a word this window regenerates), and writes only word i+k.
    const uint64_t* win = nullptr;
    int win_i0 = 0;
    int win_n = 0;

    RSB_HD uint64_t& w(int i) { return mt[(size_t)i * stride]; }

    // random.pyx:110-122
    RSB_HD void init_genrand64(uint64_t seed) {
        w(0) = seed;
        uint64_t prev = seed;
        for (int i = 1; i < RSB_MT_NN; ++i) {
            prev = 6364136223846793005ULL * (prev ^ (prev >> 62)) + (uint64_t)i;
            w(i) = prev;
        }
        mti = RSB_MT_NN;
    }

    // random.pyx:125-164 specialised to the key seed(d) builds for 0 < d < 2^64 (random.pyx:215-243):
    // d.to_bytes(8*312, 'big') => key[0..310] = 0, key[311] = d.
    RSB_HD void seed(uint64_t d) {
        init_genrand64(19650218ULL);
        unsigned int i = 1, j = 0;
        for (int k = 0; k < RSB_MT_NN; ++k) {
            uint64_t key = (j == RSB_MT_NN - 1) ? d : 0ULL;
            uint64_t prev = w(i - 1);
            w(i) = (w(i) ^ ((prev ^ (prev >> 62)) * 3935559000370003845ULL)) + key + (uint64_t)j;
            ++i;
            if (i >= RSB_MT_NN) { w(0) = w(RSB_MT_NN - 1); i = 1; }
            ++j;
            if (j >= RSB_MT_NN) j = 0;
        }
        for (int k = 0; k < RSB_MT_NN - 1; ++k) {
            uint64_t prev = w(i - 1);
            w(i) = (w(i) ^ ((prev ^ (prev >> 62)) * 2862933555777941757ULL)) - (uint64_t)i;
            ++i;
            if (i >= RSB_MT_NN) { w(0) = w(RSB_MT_NN - 1); i = 1; }
        }
        w(0) = 9223372036854775808ULL;
        mti = RSB_MT_NN;
    }

    // random.pyx:167-212 (_rand_uint64).  The reference regenerates all 312 words when the cursor runs
    // out; the recurrence only ever reads word i (old), word i+1 (old, or new word 0 for i = 311) and word
    // i+156 mod 312 (old for i < 156, already regenerated for i >= 156), so regenerating word i lazily,
    // in place, at the moment it is drawn yields the identical sequence -- with 3 loads + 1 store per draw
    // and no 312-trip refill loop for a diverged warp to serialise on.
    RSB_HD uint64_t next_u64() {
        const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, MAG = 0xB5026F5AA96619E9ULL;
        int i = (mti >= RSB_MT_NN) ? 0 : mti;
        int i1 = (i + 1 == RSB_MT_NN) ? 0 : i + 1;
        int im = (i + RSB_MT_MM >= RSB_MT_NN) ? i + RSB_MT_MM - RSB_MT_NN : i + RSB_MT_MM;
        uint64_t wi, wi1, wim;
        int k = i - win_i0;
        if (k < 0) k += RSB_MT_NN;
        if (k < win_n) {
            wi = win[k * RSB_MT_WIN_STRIDE];
            wi1 = win[(k + 1) * RSB_MT_WIN_STRIDE];
            wim = win[(k + 7) * RSB_MT_WIN_STRIDE];
        } else {
            wi = w(i); wi1 = w(i1); wim = w(im);
        }
        uint64_t x = (wi & UM) | (wi1 & LM);
        x = wim ^ (x >> 1) ^ ((x & 1ULL) ? MAG : 0ULL);
        w(i) = x;
        mti = i + 1;
        x ^= (x >> 29) & 0x5555555555555555ULL;
        x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
        x ^= (x << 37) & 0xFFF7EEE000000000ULL;
        x ^= (x >> 43);
        return x;
    }
};

// seed(d) in 312 dependent steps instead of 935.  seed(d) keys init_by_array64 with (0, ..., 0, d) (random.pyx:215-243),
// so init_genrand64(19650218) and the first 311 steps of the first mixing loop (key word 0, random.pyx:141-154) do not
// depend on d: their result is one constant table T (mt_seed_table).  What is left per stream: the 312th step of that
// loop (the only one that adds d), the 311 steps of the second loop, and mt[0] = 2^63.
RSB_HD void mt_seed_table(uint64_t* T) {
    T[0] = 19650218ULL;
    for (int i = 1; i < RSB_MT_NN; ++i) T[i] = 6364136223846793005ULL * (T[i - 1] ^ (T[i - 1] >> 62)) + (uint64_t)i;
    for (int k = 0; k < RSB_MT_NN - 1; ++k) {
        const int i = k + 1;
        T[i] = (T[i] ^ ((T[i - 1] ^ (T[i - 1] >> 62)) * 3935559000370003845ULL)) + (uint64_t)k;
    }
    T[0] = T[RSB_MT_NN - 1];     // i reached NN: mt[0] = mt[NN - 1], i = 1
}

// word 1 after the first mixing loop's last step (i = 1, j = 311: the step that adds d)
RSB_HD uint64_t mt_seed_first(const uint64_t* T, uint64_t d) {
    return (T[1] ^ ((T[0] ^ (T[0] >> 62)) * 3935559000370003845ULL)) + d + (uint64_t)(RSB_MT_NN - 1);
}
// second mixing loop, word i (2 <= i < NN) from the word before it
RSB_HD uint64_t mt_seed_step(uint64_t t_i, uint64_t prev, int i) {
    return (t_i ^ ((prev ^ (prev >> 62)) * 2862933555777941757ULL)) - (uint64_t)i;
}
// ... and its last step, back at i = 1 with mt[0] = the new mt[NN - 1]
RSB_HD uint64_t mt_seed_last(uint64_t m1, uint64_t m_last) {
    return (m1 ^ ((m_last ^ (m_last >> 62)) * 2862933555777941757ULL)) - 1ULL;
}

// the same state Mt19937_64::seed(d) leaves, through the table (serial form: tests, and the statement the warp-
// cooperative k_wf_seed follows)
RSB_HD void mt_seed_fast(const uint64_t* T, uint64_t d, uint64_t* m) {
    const uint64_t m1 = mt_seed_first(T, d);
    uint64_t prev = m1;
    for (int i = 2; i < RSB_MT_NN; ++i) {
        prev = mt_seed_step(T[i], prev, i);
        m[i] = prev;
    }
    m[1] = mt_seed_last(m1, prev);
    m[0] = 9223372036854775808ULL;
}

// Seeds a pixel's two cursors in one pass over a thread-private scratch array (local memory, L1 resident)
// instead of running the 935-step initialisation chain and the 2*spp-draw skip as dependent round trips to
// the HBM-resident state: `jitter` receives the state right after seed(d) (cursor at draw 0), `path` the
// same stream advanced by `skip` draws.  Sequence identical to Mt19937_64::seed + next_u64.
RSB_HD void mt_seed_pair(uint64_t d, int skip, uint64_t* jitter, int* jitter_mti, uint64_t* path, int* path_mti) {
    uint64_t m[RSB_MT_NN];
    Mt19937_64 g;
    g.mt = m;
    g.stride = 1;
    g.seed(d);
    for (int i = 0; i < RSB_MT_NN; ++i) jitter[i] = m[i];
    *jitter_mti = g.mti;
    for (int i = 0; i < skip; ++i) (void)g.next_u64();
    for (int i = 0; i < RSB_MT_NN; ++i) path[i] = m[i];
    *path_mti = g.mti;
}

struct Philox4x32 {
    uint32_t key[2];
    uint32_t ctr[3];    // (sub-stream, stream lo, stream hi); the block index is idx >> 1
    uint32_t idx;       // 64-bit words drawn so far: the whole state besides the key and the stream id

    RSB_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
        uint64_t p = (uint64_t)a * (uint64_t)b;
        *hi = (uint32_t)(p >> 32);
        *lo = (uint32_t)p;
    }

    RSB_HD void init(uint64_t seed, uint64_t stream, uint32_t sub) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = sub;
        ctr[1] = (uint32_t)stream;
        ctr[2] = (uint32_t)(stream >> 32);
        idx = 0;
    }

    // Philox4x32-10 (Salmon et al., SC'11) of counter (idx >> 1, ctr[0..2]); word idx & 1 of the block
    RSB_HD uint64_t next_u64() {
        uint32_t c0 = idx >> 1, c1 = ctr[0], c2 = ctr[1], c3 = ctr[2];
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, &hi0, &lo0);
            mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0;
            uint32_t n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        uint64_t w = (idx & 1u) ? (((uint64_t)c3 << 32) | (uint64_t)c2) : (((uint64_t)c1 << 32) | (uint64_t)c0);
        idx += 1;
        return w;
    }
};

enum RngMode : int32_t { RNG_MT19937_64 = 0, RNG_PHILOX = 1 };

// One stream handle used by the trace loop.
struct Rng {
    int32_t mode;
    Mt19937_64 mt;
    Philox4x32 px;

    RSB_HD uint64_t next_u64() { return mode == RNG_MT19937_64 ? mt.next_u64() : px.next_u64(); }

    // random.pyx:247-265 (uniform): (x >> 11) * 2^-53
    RSB_HD double uniform() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }

    // random.pyx:308-332 (probability)
    RSB_HD bool probability(double prob) { return uniform() < prob; }
};

}  // namespace rsb
