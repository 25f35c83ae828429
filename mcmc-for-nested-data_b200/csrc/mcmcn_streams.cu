// mcmcn_streams.cu -- the chains' HOST random streams of the start state (host code only; no kernel here).
//
// The reference runs every chain in its own process and seeds numpy's global legacy generator with the chain
// index (posteriorSampling.py:225 `numpy.random.seed(seed)`, seed = chain, :1015); the start state is then drawn
// from that stream: numpy.random.uniform per parameter (:1069-1077), numpy.random.normal per parameter and group
// under partial pooling (:738-758).  To start every chain where the reference starts it, each chain here owns the
// same stream: MT19937 seeded by init_genrand(chain), 53-bit doubles from two outputs, the polar (Marsaglia)
// normal with its cached second value -- restated from the published algorithms (Matsumoto & Nishimura 1998;
// numpy's legacy `RandomState` distributions), checked bit for bit against numpy.random.RandomState in
// tests/test_host_cpu.py.  Thousands of streams are advanced by host threads: at 16,384 chains the Python loop
// over RandomState objects took 5 s per 8,192 chains (1.7 s constructing them, 3 s drawing); this takes a
// fraction of a second.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "mcmcn_host.h"

namespace {

struct Stream {
    uint32_t key[624];
    int32_t pos;
    int32_t has_gauss;
    double gauss;
};

struct Streams {
    int64_t n;
    std::vector<Stream> s;
};

void seed_stream(Stream& st, uint32_t seed) {          // init_genrand
    for (int i = 0; i < 624; ++i) {
        st.key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    st.pos = 624;
    st.has_gauss = 0;
    st.gauss = 0.0;
}

void refill(Stream& st) {                              // the 624-word block of the recurrence
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    uint32_t* k = st.key;
    int i = 0;
    for (; i < 624 - 397; ++i) {
        const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; i < 623; ++i) {
        const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
        k[i] = k[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    const uint32_t y = (k[623] & UPPER) | (k[0] & LOWER);
    k[623] = k[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    st.pos = 0;
}

inline uint32_t next32(Stream& st) {
    if (st.pos == 624) refill(st);
    uint32_t y = st.key[st.pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

inline double next_double(Stream& st) {                // genrand_res53
    const uint32_t a = next32(st) >> 5, b = next32(st) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

inline double next_gauss(Stream& st) {                 // polar method; the second value of a pair waits in the stream
    if (st.has_gauss) {
        const double v = st.gauss;
        st.has_gauss = 0;
        st.gauss = 0.0;
        return v;
    }
    double x1, x2, r2;
    do {
        x1 = 2.0 * next_double(st) - 1.0;
        x2 = 2.0 * next_double(st) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = std::sqrt(-2.0 * std::log(r2) / r2);
    st.gauss = f * x1;
    st.has_gauss = 1;
    return f * x2;
}

template <typename F>
void over_streams(int64_t n, int threads, F body) {    // body(j) for j in [0, n), contiguous ranges per thread
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    if (nt == 1) {
        for (int64_t j = 0; j < n; ++j) body(j);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        pool.emplace_back([=]() {
            for (int64_t j = lo; j < hi; ++j) body(j);
        });
    }
    for (auto& th : pool) th.join();
}

bool listed_ok(const Streams* h, const int64_t* which, int64_t n_which) {
    if (!h || n_which < 0 || (n_which > 0 && !which)) return false;
    for (int64_t j = 0; j < n_which; ++j)
        if (which[j] < 0 || which[j] >= h->n) return false;
    return true;
}

}  // namespace

extern "C" {

int mcmcn_streams_create(int64_t n, int64_t seed0, int32_t threads, void** out_handle) {
    if (n < 0 || seed0 < 0 || seed0 + n - 1 > 0xffffffffLL || !out_handle) {
        mcmcn::set_error("mcmcn_streams_create: n >= 0 and seeds in [0, 2^32) required");
        return MCMCN_ERR_INVALID;
    }
    Streams* h = new (std::nothrow) Streams();
    if (!h) {
        mcmcn::set_error("mcmcn_streams_create: out of memory");
        return MCMCN_ERR_INVALID;
    }
    h->n = n;
    h->s.resize((size_t)n);
    over_streams(n, threads, [&](int64_t j) { seed_stream(h->s[(size_t)j], (uint32_t)(seed0 + j)); });
    *out_handle = h;
    return MCMCN_OK;
}

int mcmcn_streams_free(void* handle) {
    delete static_cast<Streams*>(handle);
    return MCMCN_OK;
}

int mcmcn_streams_uniform(void* handle, const int64_t* which, int64_t n_which, int32_t count, const double* low,
                          const double* high, double* out, int32_t threads) {
    Streams* h = static_cast<Streams*>(handle);
    if (!listed_ok(h, which, n_which) || count < 0 || !low || !high || !out) {
        mcmcn::set_error("mcmcn_streams_uniform: bad arguments");
        return MCMCN_ERR_INVALID;
    }
    over_streams(n_which, threads, [&](int64_t j) {
        Stream& st = h->s[(size_t)which[j]];
        double* o = out + (size_t)j * (size_t)count;
        for (int i = 0; i < count; ++i) {
            const double range = high[i] - low[i];
            o[i] = low[i] + range * next_double(st);
        }
    });
    return MCMCN_OK;
}

int mcmcn_streams_normal(void* handle, const int64_t* which, int64_t n_which, const int64_t* out_off, double* out,
                         int32_t threads) {
    Streams* h = static_cast<Streams*>(handle);
    if (!listed_ok(h, which, n_which) || !out_off || !out) {
        mcmcn::set_error("mcmcn_streams_normal: bad arguments");
        return MCMCN_ERR_INVALID;
    }
    for (int64_t j = 0; j < n_which; ++j)
        if (out_off[j + 1] < out_off[j]) {
            mcmcn::set_error("mcmcn_streams_normal: offsets must not decrease");
            return MCMCN_ERR_INVALID;
        }
    over_streams(n_which, threads, [&](int64_t j) {
        Stream& st = h->s[(size_t)which[j]];
        for (int64_t i = out_off[j]; i < out_off[j + 1]; ++i) out[i] = next_gauss(st);
    });
    return MCMCN_OK;
}

int mcmcn_streams_get_state(void* handle, int64_t i, uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* gauss) {
    Streams* h = static_cast<Streams*>(handle);
    if (!h || i < 0 || i >= h->n || !key624 || !pos || !has_gauss || !gauss) {
        mcmcn::set_error("mcmcn_streams_get_state: bad arguments");
        return MCMCN_ERR_INVALID;
    }
    const Stream& st = h->s[(size_t)i];
    std::memcpy(key624, st.key, sizeof(st.key));
    *pos = st.pos;
    *has_gauss = st.has_gauss;
    *gauss = st.gauss;
    return MCMCN_OK;
}

int mcmcn_streams_set_state(void* handle, int64_t i, const uint32_t* key624, int32_t pos, int32_t has_gauss, double gauss) {
    Streams* h = static_cast<Streams*>(handle);
    if (!h || i < 0 || i >= h->n || !key624 || pos < 0 || pos > 624) {
        mcmcn::set_error("mcmcn_streams_set_state: bad arguments");
        return MCMCN_ERR_INVALID;
    }
    Stream& st = h->s[(size_t)i];
    std::memcpy(st.key, key624, sizeof(st.key));
    st.pos = pos;
    st.has_gauss = has_gauss ? 1 : 0;
    st.gauss = gauss;
    return MCMCN_OK;
}

}  // extern "C"
