/* AF-RNG in host C -- ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Independent restatement of the normative spec in oracle/afrng.py (Philox4x32-10,
 * 53-bit uniforms, fdlibm-scheme log/exp, polar normal, chunked-product Poisson,
 * the reference's general_sampler mapping, samplers/common_helpers.py:49-89).
 * It exists so that the Python oracle (oracle/des_port.py) pays about what the
 * reference pays numpy for a draw instead of ~15 us of pure-Python Philox.
 * tests/test_afrng.py checks it bit-for-bit against afrng.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (see oracle/Makefile)
 * -ffp-contract=off is REQUIRED: no fused multiply-add may be formed.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define P_GEN 0u
#define P_EDGE 1u
#define P_SERVER 2u

enum { D_POISSON = 0, D_NORMAL = 1, D_LOG_NORMAL = 2, D_EXPONENTIAL = 3, D_UNIFORM = 4 };

void afrng_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t hi, uint32_t lo) {
    return (double)(((uint64_t)(hi >> 5) << 26) + (lo >> 6)) * (1.0 / 9007199254740992.0);
}
static inline double s32(uint32_t w) { return ((double)w + 0.5) * (1.0 / 2147483648.0) - 1.0; }

static inline uint64_t d2u(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }
static inline double u2d(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }

double afrng_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
        Lg7 = 1.479819860511658591e-01;
    uint64_t b = d2u(x);
    int32_t hx = (int32_t)(b >> 32);
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000FFFFF;
    int32_t i = (hx + 0x95F64) & 0x100000;
    b = ((uint64_t)(uint32_t)(hx | (i ^ 0x3FF00000)) << 32) | (b & 0xFFFFFFFFu);
    k += i >> 20;
    double f = u2d(b) - 1.0;
    double dk = (double)k;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double r = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + r) + dk * ln2_lo)) - f);
}

double afrng_exp(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
        P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
        P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    int32_t k = (x >= 0.0) ? (int32_t)(invln2 * x + 0.5) : (int32_t)(invln2 * x - 0.5);
    double dk = (double)k;
    double hi = x - dk * ln2_hi;
    double lo = dk * ln2_lo;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52));
}

/* ------------------------------------------------------------------ sources */
typedef struct {
    uint32_t key[2];
    uint32_t c2, c3;
    uint32_t idx;     /* GEN: unused; request draw: request id */
    uint32_t tag;     /* purpose<<24 | hop<<8 */
    uint32_t pos;     /* GEN: index of next uniform; request: next free half-block */
    int is_gen;
    uint32_t cached_block;
    int have_block;
    uint32_t w[4];
} af_src;

static void src_block(af_src *s, uint32_t b) {
    if (s->have_block && s->cached_block == b) return;
    uint32_t ctr[4];
    if (s->is_gen) { ctr[0] = b; ctr[1] = P_GEN << 24; }
    else { ctr[0] = s->idx; ctr[1] = s->tag | (b & 0xFFu); }
    ctr[2] = s->c2; ctr[3] = s->c3;
    afrng_philox(ctr, s->key, s->w);
    s->cached_block = b; s->have_block = 1;
}

static double src_next53(af_src *s) {
    uint32_t p = s->pos;
    if (s->is_gen) {
        s->pos = p + 1;
        src_block(s, p >> 1);
        return (p & 1u) ? u53(s->w[2], s->w[3]) : u53(s->w[0], s->w[1]);
    }
    s->pos = p + 2;
    src_block(s, p >> 2);
    return u53(s->w[p & 3u], s->w[(p & 3u) + 1]);
}

static void src_pair32(af_src *s, double *v1, double *v2) {
    if (s->is_gen) {
        double a = src_next53(s), b = src_next53(s);
        *v1 = 2.0 * a - 1.0; *v2 = 2.0 * b - 1.0;
        return;
    }
    uint32_t p = s->pos;
    s->pos = p + 2;
    src_block(s, p >> 2);
    *v1 = s32(s->w[p & 3u]); *v2 = s32(s->w[(p & 3u) + 1]);
}

static double std_normal(af_src *s) {
    for (;;) {
        double v1, v2;
        src_pair32(s, &v1, &v2);
        double q = v1 * v1 + v2 * v2;
        if (q > 0.0 && q < 1.0) return v1 * sqrt(-2.0 * afrng_log(q) / q);
    }
}

static int64_t poisson(double lam, af_src *s) {
    int64_t n = 0;
    double rem = lam;
    while (rem > 0.0) {
        double c = rem < 256.0 ? rem : 256.0;
        rem = rem - c;
        double limit = afrng_exp(-c);
        double p = 1.0;
        for (;;) {
            p = p * (1.0 - src_next53(s));
            if (p <= limit) break;
            ++n;
        }
    }
    return n;
}

static double sample_rv(int dist, double mean, double sigma, af_src *s) {
    switch (dist) {
    case D_UNIFORM: return src_next53(s);
    case D_POISSON: return (double)poisson(mean, s);
    case D_EXPONENTIAL: return mean * -afrng_log(1.0 - src_next53(s));
    case D_NORMAL: { double v = mean + sigma * std_normal(s); return v > 0.0 ? v : 0.0; }
    case D_LOG_NORMAL: return afrng_exp(mean + sigma * std_normal(s));
    default: return NAN;
    }
}

static void src_init(af_src *s, uint64_t seed, uint64_t replica) {
    memset(s, 0, sizeof *s);
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->c2 = (uint32_t)replica; s->c3 = (uint32_t)(replica >> 32);
}

/* ------------------------------------------------------------------ exports */
/* generator stream: caller keeps `pos` (index of the next uniform) */
double afrng_gen_uniform(uint64_t seed, uint64_t replica, uint32_t *pos) {
    af_src s; src_init(&s, seed, replica); s.is_gen = 1; s.pos = *pos;
    double u = src_next53(&s); *pos = s.pos; return u;
}
int64_t afrng_gen_poisson(uint64_t seed, uint64_t replica, uint32_t *pos, double lam) {
    af_src s; src_init(&s, seed, replica); s.is_gen = 1; s.pos = *pos;
    int64_t n = poisson(lam, &s); *pos = s.pos; return n;
}
double afrng_gen_normal(uint64_t seed, uint64_t replica, uint32_t *pos, double mean, double sigma) {
    af_src s; src_init(&s, seed, replica); s.is_gen = 1; s.pos = *pos;
    double v = mean + sigma * std_normal(&s); *pos = s.pos; return v;
}
/* one edge traversal: dropout uniform + latency variate */
void afrng_edge(uint64_t seed, uint64_t replica, uint32_t rid, uint32_t hop, int dist,
                double mean, double sigma, double *u_drop, double *latency) {
    af_src s; src_init(&s, seed, replica);
    s.idx = rid; s.tag = (P_EDGE << 24) | ((hop & 0xFFFFu) << 8); s.pos = 2;
    src_block(&s, 0);
    *u_drop = u53(s.w[0], s.w[1]);
    *latency = sample_rv(dist, mean, sigma, &s);
}
uint32_t afrng_endpoint(uint64_t seed, uint64_t replica, uint32_t rid, uint32_t hop, uint32_t n) {
    af_src s; src_init(&s, seed, replica);
    s.idx = rid; s.tag = (P_SERVER << 24) | ((hop & 0xFFFFu) << 8);
    src_block(&s, 0);
    return (uint32_t)(((uint64_t)s.w[0] * n) >> 32);
}
