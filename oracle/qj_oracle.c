/*
 * qj_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C + OpenMP) of the state-vector hot path of qibojit's numba
 * backend.  It is the checker for the CUDA path and the timed `cpu_baseline` /
 * `--impl reference` arm of bench.py.  Nothing under qibojit_b200/ may import, link
 * or execute it.
 *
 * Pinning: the tests/golden/ fixtures were produced by importing the reference's own numba
 * kernels (tests/golden/make_golden.py) and tests/test_oracle.py checks this file
 * against them, plus the reference's sampler golden vector tests/test_ops.py:251-256.
 * `qjo_probabilities_*` restates qibo code that is not in /root/reference: parity
 * unpinned for that one function.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { QJO_OP_GATE = 0, QJO_OP_X = 1, QJO_OP_Y = 2, QJO_OP_Z = 3, QJO_OP_ZPOW = 4,
       QJO_OP_SWAP = 5, QJO_OP_FSIM = 6 };

/* gates.py:6-12 -- insert a 1 at every active position (ascending). */
static inline int64_t mc_index(int64_t g, const int32_t *qubits, int nq)
{
    int64_t i = g;
    for (int j = 0; j < nq; j++) {
        const int n = qubits[j];
        const int64_t k = (int64_t)1 << n;
        i = ((i >> n) << (n + 1)) + (i & (k - 1)) + k;
    }
    return i;
}

/* gates.py:258-262 */
static inline int64_t mt_index(int64_t i, const int64_t *targets, int nt)
{
    int64_t t = 0;
    for (int u = 0; u < nt; u++) t += ((i >> u) & 1) * targets[u];
    return t;
}

/* ops.py:34-40 */
static inline int64_t collapse_index(int64_t g, int64_t h, const int32_t *qubits, int nq)
{
    int64_t i = g;
    for (int j = 0; j < nq; j++) {
        const int n = qubits[j];
        const int64_t k = (int64_t)1 << n;
        i = ((i >> n) << (n + 1)) + (i & (k - 1)) + ((h >> j) & 1) * k;
    }
    return i;
}

/* ---- MT19937 as numba / numpy legacy RandomState use it ----------------------
 * seeding: numba/_random.c:59-75; twist: _random.c:37-56; tempering + double:
 * numba/cpython/randomimpl.py:109-147; bounded ints: randomimpl.py:149-196,454-523. */
typedef struct { uint32_t mt[624]; int idx; } mt_state;

static void mt_seed(mt_state *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static void mt_twist(mt_state *s)
{
    uint32_t *mt = s->mt;
    for (int i = 0; i < 624; i++) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        uint32_t v = mt[(i + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        mt[i] = v;
    }
    s->idx = 0;
}

static inline uint32_t mt_u32(mt_state *s)
{
    if (s->idx >= 624) mt_twist(s);
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static inline double mt_double(mt_state *s)
{
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return ((double)b + (double)a * 67108864.0) / 9007199254740992.0;
}

/* np.random.randint(0, n) inside @njit with int64 arguments */
static inline int64_t mt_randint(mt_state *s, int64_t n)
{
    if (n == 1) return 0;
    int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
    for (;;) {
        int64_t r;
        if (nbits <= 32) {
            uint32_t mask = 0xffffffffu >> (32 - nbits);
            r = (int64_t)(mt_u32(s) & mask);
        } else {
            uint32_t mask = 0xffffffffu >> (64 - nbits);
            uint64_t high = mt_u32(s) & mask;
            uint64_t low = mt_u32(s);
            r = (int64_t)(low + (high << 32));
        }
        if (r < n) return r;
    }
}

/* exported for tests: fills out[0..count) with the double stream after seeding */
void qjo_mt_doubles(int64_t seed, int count, double *out)
{
    mt_state s; mt_seed(&s, (uint32_t)seed);
    for (int i = 0; i < count; i++) out[i] = mt_double(&s);
}
void qjo_mt_randints(int64_t seed, int64_t n, int count, int64_t *out)
{
    mt_state s; mt_seed(&s, (uint32_t)seed);
    for (int i = 0; i < count; i++) out[i] = mt_randint(&s, n);
}

#define qjo_hypot_c64 hypotf
#define qjo_hypot_c128 hypot

#define REAL float
#define CPLX cplx_c64
#define SUF(name) name##_c64
#include "qj_oracle_kernels.inc"
#undef REAL
#undef CPLX
#undef SUF

#define REAL double
#define CPLX cplx_c128
#define SUF(name) name##_c128
#include "qj_oracle_kernels.inc"
#undef REAL
#undef CPLX
#undef SUF

int qjo_max_threads(void);
#ifdef _OPENMP
#include <omp.h>
int qjo_max_threads(void) { return omp_get_max_threads(); }
void qjo_set_threads(int n) { omp_set_num_threads(n); }
#else
int qjo_max_threads(void) { return 1; }
void qjo_set_threads(int n) { (void)n; }
#endif
