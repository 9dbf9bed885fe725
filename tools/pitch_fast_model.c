/*
 * pitch_fast_model.c -- CPU model of the DECISION LOGIC of the fast pitch kernel (nnnoiseless_b200/csrc/pitch.cu),
 * checked against the oracle's order-exact pitch path on millions of frames before any GPU time is spent.
 *
 * TEST INFRASTRUCTURE (like oracle/): it #includes the oracle source to reach its static functions.
 *
 * The fast kernel computes the two big families of dot products -- the 147-lag coarse cross-correlation
 * (src/pitch.rs:82, 296-363) and the sub-harmonic inner products of remove_doubling (src/pitch.rs:152-168) --
 * with FMA in an arbitrary order, and proves per stream that every DECISION the reference takes from them
 * is unchanged.  For any order / any mix of fused and unfused roundings
 *      |computed - exact| <= gamma_n * sum |x_j y_j| <= gamma_n * ||x|| ||y||,  gamma_n = n u / (1 - n u), u = 2^-24,
 * so a fast value and the reference's value differ by at most 2 gamma_n ||x|| ||y||.
 *
 * Coarse search (find_best_pitch over 147 lags): score r_i = c_i^2 / y_i for c_i > 0.
 *   approximate top two F1, F2; T0 = min(lo_F1, lo_F2); candidate set C = { j : hi_j (1+eta) >= T0 }.
 *   |C| = 2 and lo_F1 > hi_F2 (1+eta)           -> (best, second) = (F1, F2), nothing exact needed.
 *   otherwise (|C| <= CMAX): exact c_j for j in C, every one must beat M = max_{j not in C} hi_j robustly,
 *                            then the reference's sequential selection restricted to C (ascending lag).
 *   Lemma: if every element of a set "top" (|top| >= 2) robustly beats every other
 *   element, the reference's scan ends with the same (best, second) as the scan restricted to top.
 * remove_doubling: g1 > thresh decided with |g1^ - thresh^| > dg1 + 0.9 dg0 + 1e-6, else the stream is flagged.
 * Everything that becomes STATE or OUTPUT (last_gain, the +-1 refinement) is recomputed order-exact.
 * Flagged streams take the exact kernel.
 *
 *   gcc -O2 -march=native -ffp-contract=off -fopenmp -o /tmp/pfm tools/pitch_fast_model.c -lm && /tmp/pfm [streams] [frames]
 */
#include "../oracle/nno_oracle.c"

#include <stdio.h>

#define NL4 147
#define N4 240
#define CMAX 8
static const float KAPPA4 = 3.1e-5f; /* 2 gamma_256 (coarse sums: 240 terms) */
static const float KAPPA2 = 6.0e-5f; /* 2 gamma_496 (480-term inner products) */
static const float ETA = 1e-5f;

typedef struct {
    long frames, flagged, flag_nonfinite, flag_cmax, flag_weakcand, flag_rd, flag_sign;
    long exact_lags, tie12, mism_period, mism_gain, csum;
} stats_t;

static float fma_dot(const float *x, const float *y, int n) { /* sequential fused chain: the GPU's coarse order */
    float c = 0.0f;
    for (int j = 0; j < n; j++) c = fmaf(x[j], y[j], c);
    return c;
}
static float chunk_dot(const float *x, const float *y) { /* 32 lanes x 15 samples, tree-reduced: the GPU's rd order */
    float part[32];
    for (int l = 0; l < 32; l++) {
        float c = 0.0f;
        for (int j = 0; j < 15; j++) c = fmaf(x[15 * l + j], y[15 * l + j], c);
        part[l] = c;
    }
    for (int o = 16; o >= 1; o >>= 1)
        for (int l = 0; l < o; l++) part[l] += part[l + o];
    return part[0];
}

/* returns 1 if flagged; else *best/*second as the reference's find_best_pitch(xcorr, y_lp4, 240) */
static int coarse_fast(const float *x4, const float *y4, int *best, int *second, stats_t *st) {
    float ch[NL4], yn[NL4];
    float ex = 0.0f, ytot = 0.0f;
    for (int j = 0; j < N4; j++) ex = fmaf(x4[j], x4[j], ex);
    for (int j = 0; j < N4 + NL4; j++) ytot = fmaf(y4[j], y4[j], ytot);
    /* exact running energy, as the reference (src/pitch.rs:379-382, 401-402) */
    float y = 1.0f;
    for (int j = 0; j < N4; j++) y += y4[j] * y4[j];
    for (int i = 0; i < NL4; i++) {
        yn[i] = y;
        y += y4[i + N4] * y4[i + N4] - y4[i] * y4[i];
        y = fmaxf(y, 1.0f);
    }
    for (int i = 0; i < NL4; i++) ch[i] = fma_dot(x4, y4 + i, N4);
    const float delta = KAPPA4 * sqrtf(ex * ytot) * 1.001f;
    if (!(delta < 1e30f)) { st->flag_nonfinite++; return 1; }   /* inf / nan */
    int anynz = 0;
    for (int j = 0; j < N4 + NL4; j++) anynz |= (y4[j] != 0.0f);
    for (int j = 0; j < N4; j++) anynz |= (x4[j] != 0.0f);
    /* exactly silent history: every c_i is exactly 0, the reference keeps its initial (0, 1) */
    if (!anynz) { *best = 0; *second = 1; return 0; }
    /* Outside the range where find_best_pitch's own products c^2 y stay finite and normal (c^2 <= ex ytot, 1 <= y <= ytot + 1)
     * the reference selects on signs alone (underflow) or gets stuck on inf > inf (overflow): parity means reproducing that,
     * so those streams take the exact path.  Inside the range every score below is a finite, normal quotient. */
    if (!(ex > 1e-18f) || !(ytot > 1e-18f) || !(ex * ytot * (ytot + 1.0f) < 1e37f)) { st->flag_nonfinite++; return 1; }
    /* approximate top two by (rounded) score ch^2 / yn (ch > 0): any two distinct lags keep the certificate sound */
    int f1 = -1, f2 = -1;
    float r1 = 0.f, r2 = 0.f;
    for (int i = 0; i < NL4; i++) {
        if (!(ch[i] > 0.0f)) continue;
        float r = ch[i] * ch[i] / yn[i];
        if (r > r1) { f2 = f1; r2 = r1; f1 = i; r1 = r; }
        else if (r > r2) { f2 = i; r2 = r; }
    }
    if (f1 < 0 || f2 < 0) { st->flag_sign++; return 1; }
    const float a1 = ch[f1] - delta, a2 = ch[f2] - delta;
    if (!(a1 > 0.0f) || !(a2 > 0.0f)) { st->flag_sign++; return 1; }
    if (!(a1 * a1 > 1e-20f) || !(a2 * a2 > 1e-20f)) { st->flag_nonfinite++; return 1; }
    const float ETA1 = 1.0f + ETA;
    const float lo1 = a1 * a1 / yn[f1], lo2 = a2 * a2 / yn[f2];
    const float b1 = ch[f1] + delta, b2 = ch[f2] + delta;
    const float hi1 = b1 * b1 / yn[f1], hi2 = b2 * b2 / yn[f2];
    const float t0s = fminf(lo1, lo2);
    int cand[CMAX], nc = 0;
    float hmax = 0.0f; /* M = max hi over non-candidates */
    for (int j = 0; j < NL4; j++) {
        float b = fmaxf(ch[j] + delta, 0.0f);
        float hs = b * b / yn[j];
        int in_c = (j == f1 || j == f2) || (hs * ETA1 >= t0s);
        if (in_c) {
            if (nc == CMAX) { st->flag_cmax++; return 1; }
            cand[nc++] = j;
        } else if (hs > hmax) hmax = hs;
    }
    const float mbound = hmax * ETA1;
    st->csum += nc;
    if (nc == 2) {
        if (lo1 > hi2 * ETA1) { *best = f1; *second = f2; return 0; }
        if (lo2 > hi1 * ETA1) { *best = f2; *second = f1; return 0; }
        st->tie12++;
    }
    /* exact values for the candidates, each must beat every non-candidate robustly */
    float ce[CMAX];
    for (int k = 0; k < nc; k++) {
        float c = 0.0f;
        for (int j = 0; j < N4; j++) c += x4[j] * y4[cand[k] + j];
        ce[k] = c;
        st->exact_lags++;
        if (!(c > 0.0f) || !(c * c / yn[cand[k]] > mbound * ETA1)) { st->flag_weakcand++; return 1; }
    }
    /* the reference's selection restricted to C (ascending lag; cand[] is ascending) */
    float best_num = -1.0f, second_num = -1.0f, best_den = 0.0f, second_den = 0.0f;
    int bp = 0, sp = 1;
    for (int k = 0; k < nc; k++) {
        float corr = ce[k], ysq = yn[cand[k]];
        if (corr > 0.0f) {
            float num = corr * corr;
            if (num * second_den > second_num * ysq) {
                if (num * best_den > best_num * ysq) { second_num = best_num; second_den = best_den; sp = bp; best_num = num; best_den = ysq; bp = cand[k]; }
                else { second_num = num; second_den = ysq; sp = cand[k]; }
            }
        }
    }
    *best = bp; *second = sp;
    return 0;
}

/* fast pitch_search: coarse fast, fine exact.  returns -1 if flagged */
static int pitch_search_fast(nno_state *s, stats_t *st) {
    const float *x_lp = s->pitch_buf + PITCH_MAX_PERIOD / 2;
    const float *y = s->pitch_buf;
    const int len = PITCH_FRAME_SIZE, max_pitch = PITCH_MAX_PERIOD - 3 * PITCH_MIN_PERIOD;
    float x_lp4[240], y_lp4[387], xcorr[294];
    const int n4 = len / 4, ny4 = len / 4 + max_pitch / 4;
    for (int j = 0; j < n4; j++) x_lp4[j] = x_lp[2 * j];
    for (int j = 0; j < ny4; j++) y_lp4[j] = y[2 * j];
    int best, second;
    if (coarse_fast(x_lp4, y_lp4, &best, &second, st)) return -1;
    for (int i = 0; i < max_pitch / 2; i++) {
        xcorr[i] = 0.0f;
        if (abs(i - 2 * best) > 2 && abs(i - 2 * second) > 2) continue;
        xcorr[i] = fmaxf(inner_prod(x_lp, y + i, len / 2), -1.0f);
    }
    int dummy;
    find_best_pitch(xcorr, max_pitch / 2, y, len / 2, &best, &dummy);
    int offset = 0;
    if (best > 0 && best < max_pitch / 2 - 1) {
        float a = xcorr[best - 1], b = xcorr[best], c = xcorr[best + 1];
        if (c - a > 0.7f * (b - a)) offset = 1;
        else if (a - c > 0.7f * (b - c)) offset = -1;
    }
    return 2 * best - offset;
}

/* fast remove_doubling: decisions from FMA inner products with margins, state/outputs exact.  returns -1 if flagged */
static int remove_doubling_fast(nno_state *s, int pitch_idx, float *gain_out, stats_t *st) {
    const float *x = s->pitch_buf;
    const int min_period = PITCH_MIN_PERIOD / 2, max_period = PITCH_MAX_PERIOD / 2, n = PITCH_FRAME_SIZE / 2;
    int t0 = pitch_idx / 2;
    if (t0 > max_period - 1) t0 = max_period - 1;
    const int prev_period = s->last_period / 2;
    float *yy_lookup = s->scratch;
    const float xx = inner_prod(x + max_period, x + max_period, n); /* exact */
    float plow = 0.0f;
    for (int j = 0; j < max_period; j++) plow = fmaf(x[j], x[j], plow);
    const float ptot = plow + xx;
    yy_lookup[0] = xx;
    float yy = xx;
    for (int i = 1; i <= max_period; i++) {
        yy += x[max_period - i] * x[max_period - i] - x[max_period + n - i] * x[max_period + n - i];
        yy_lookup[i] = fmaxf(yy, 0.0f);
    }
    const float dip = KAPPA2 * sqrtf(xx * ptot) * 1.001f;
    if (!(dip < 1e30f)) { st->flag_nonfinite++; return -1; }
    const float xy0 = chunk_dot(x + max_period, x + max_period - t0);
    const float inv0 = 1.0f / sqrtf(1.0f + xx * yy_lookup[t0]);
    const float g0 = xy0 * inv0, dg0 = dip * inv0;
    int t = t0, ksel = 1, t1b_sel = 0;
    for (int k = 2; k <= 15; k++) {
        int t1 = (2 * t0 + k) / (2 * k);
        if (t1 < min_period) break;
        int t1b;
        if (k == 2) t1b = (t1 + t0 > max_period) ? t0 : t0 + t1;
        else t1b = (2 * SECOND_CHECK[k] * t0 + k) / (2 * k);
        float xy = 0.5f * (chunk_dot(x + max_period, x + max_period - t1) + chunk_dot(x + max_period, x + max_period - t1b));
        float yyv = (yy_lookup[t1] + yy_lookup[t1b]) / 2.0f;
        float inv = 1.0f / sqrtf(1.0f + xx * yyv);
        float g1 = xy * inv, dg1 = dip * inv;
        float cont;
        int d = abs(t1 - prev_period);
        if (d <= 1) cont = s->last_gain;
        else if (d <= 2 && 5 * k * k < t0) cont = s->last_gain / 2.0f;
        else cont = 0.0f;
        float thresh;
        if (t1 < 3 * min_period) thresh = fmaxf(0.85f * g0 - cont, 0.4f);
        else if (t1 < 2 * min_period) thresh = fmaxf(0.9f * g0 - cont, 0.5f);
        else thresh = fmaxf(0.7f * g0 - cont, 0.3f);
        if (!(fabsf(g1 - thresh) > dg1 + 0.9f * dg0 + 1e-6f)) { st->flag_rd++; return -1; }
        if (g1 > thresh) { t = t1; ksel = k; t1b_sel = t1b; }
    }
    /* exact: xcorr at t-1, t, t+1 (the +-1 refinement) and, for k >= 2, at t1b; then best_xy, best_yy, g, pg */
    float xc[3];
    for (int k = 0; k < 3; k++) xc[k] = inner_prod(x + max_period, x + max_period - (t + k - 1), n);
    float best_xy, best_yy;
    if (ksel == 1) { best_xy = xc[1]; best_yy = yy_lookup[t0]; }
    else {
        float xy2 = inner_prod(x + max_period, x + max_period - t1b_sel, n);
        best_xy = (xc[1] + xy2) / 2.0f;
        best_yy = (yy_lookup[t] + yy_lookup[t1b_sel]) / 2.0f;
    }
    const float g = pitch_gain(best_xy, xx, best_yy);
    best_xy = fmaxf(best_xy, 0.0f);
    float pg = (best_yy <= best_xy) ? 1.0f : best_xy / (best_yy + 1.0f);
    int offset = 0;
    if (xc[2] - xc[0] > 0.7f * (xc[1] - xc[0])) offset = 1;
    else if (xc[0] - xc[2] > 0.7f * (xc[1] - xc[2])) offset = -1;
    pg = fminf(pg, g);
    int tf = 2 * t + offset;
    if (tf < PITCH_MIN_PERIOD) tf = PITCH_MIN_PERIOD;
    *gain_out = pg;
    return tf;
}

/* ---- signal families ---------------------------------------------------------------------------------- */
static _Thread_local unsigned long long rng_s;
static double urand(void) {
    rng_s = rng_s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(rng_s >> 11) / 9007199254740992.0;
}
static double nrand(void) {
    double u1 = urand(), u2 = urand();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

int main(int argc, char **argv) {
    int n_streams = argc > 1 ? atoi(argv[1]) : 2000, n_frames = argc > 2 ? atoi(argv[2]) : 100;
    const char *rawpath = argc > 3 ? argv[3] : "tests/golden/testing.raw";
    short *raw = NULL;
    long nraw = 0;
    FILE *f = fopen(rawpath, "rb");
    if (f) {
        fseek(f, 0, SEEK_END);
        nraw = ftell(f) / 2;
        fseek(f, 0, SEEK_SET);
        raw = (short *)malloc((size_t)nraw * 2);
        if (fread(raw, 2, (size_t)nraw, f) != (size_t)nraw) nraw = 0;
        fclose(f);
    }
    ensure_tables();
    stats_t tot = {0};
#pragma omp parallel
    {
        stats_t st = {0};
#pragma omp for schedule(dynamic, 4)
        for (int sidx = 0; sidx < n_streams; sidx++) {
            unsigned long long my = 0x9E3779B97F4A7C15ULL * (unsigned long long)(sidx + 1);
            nno_state *S = nno_state_new(NULL);
            const int fam = sidx % 4;
            /* per-stream parameters */
            rng_s = my;
            double f0 = 100.0 * pow(40.0, urand()), A = 1000.0 + 11000.0 * urand(), sg = 100.0 + 2900.0 * urand(), ph = 6.283185307 * urand();
            double vib = 0.002 + 0.02 * urand(), vrate = 3.0 + 5.0 * urand();
            int nh = 2 + (int)(urand() * 10);
            double gain = 0.05 + 1.5 * urand();
            long roff = nraw ? (long)(urand() * nraw) : 0;
            if (fam == 3) sg = (sidx % 8 == 3) ? 1.0 + 20.0 * urand() : sg * 0.1; /* nearly pure tones: the near-tie stress case */
            /* amplitude regimes outside the int16 range the reference documents: underflow / overflow of the certificate */
            static const double scales[8] = {1.0, 1.0, 1.0, 1e4, 3.0, 1e-9, 3e-5, 300.0};
            const double ascale = scales[(sidx / 4) % 8];
            unsigned long long rs = rng_s;
            float in[FRAME_SIZE], hp[FRAME_SIZE];
            double phase = ph;
            for (int fr = 0; fr < n_frames; fr++) {
                rng_s = rs;
                for (int i = 0; i < FRAME_SIZE; i++) {
                    long nn = (long)fr * FRAME_SIZE + i;
                    double v;
                    if (fam == 0 || fam == 3) v = A * sin(6.283185307179586 * f0 * nn / 48000.0 + ph) + sg * nrand();
                    else if (fam == 1) { /* harmonic stack with vibrato */
                        double fi = f0 * 0.25 * (1.0 + vib * sin(6.283185307 * vrate * nn / 48000.0));
                        if (fi < 60) fi = 60;
                        phase += 6.283185307179586 * fi / 48000.0;
                        v = 0;
                        for (int h = 1; h <= nh; h++) v += A / h * sin(h * phase);
                        v += sg * 0.3 * nrand();
                    } else { /* looped speech fixture with gain, plus a little noise */
                        v = nraw ? gain * raw[(roff + nn) % nraw] + 0.02 * sg * nrand() : sg * nrand();
                        /* occasional digital silence */
                        if (((fr / 13) % 5) == 4 || (sidx % 16 == 2 && fr >= 20 && fr < 80)) v = 0;
                    }
                    v = rint(v);
                    if (v > 32767) v = 32767;
                    if (v < -32768) v = -32768;
                    in[i] = (float)(v * ascale);
                }
                rs = rng_s;
                memmove(S->input_mem, S->input_mem + FRAME_SIZE, (PITCH_BUF_SIZE - FRAME_SIZE) * sizeof(float));
                biquad_hp(hp, S->mem_hp_x, in, FRAME_SIZE);
                memcpy(S->input_mem + (PITCH_BUF_SIZE - FRAME_SIZE), hp, sizeof hp);
                /* exact */
                pitch_downsample(S->input_mem, S->pitch_buf);
                const int lp = S->last_period;
                const float lg = S->last_gain;
                int idx = PITCH_MAX_PERIOD - pitch_search(S);
                float ge;
                int pe = remove_doubling(S, idx, &ge);
                /* fast, from the same prior state */
                S->last_period = lp;
                S->last_gain = lg;
                st.frames++;
                int flagged = 0;
                int is = pitch_search_fast(S, &st);
                if (is == -1) flagged = 1;
                else {
                    float gf;
                    int pf = remove_doubling_fast(S, PITCH_MAX_PERIOD - is, &gf, &st);
                    if (pf == -1) flagged = 1;
                    else {
                        if (pf != pe) { st.mism_period++; if (st.mism_period < 4) fprintf(stderr, "mismatch stream %d (scale %g fam %d) frame %d: fast %d exact %d\n", sidx, ascale, fam, fr, pf, pe); }
                        if (memcmp(&gf, &ge, 4) != 0) st.mism_gain++;
                    }
                }
                st.flagged += flagged;
                S->last_period = pe;
                S->last_gain = ge;
            }
            nno_state_free(S);
        }
#pragma omp critical
        {
            long *a = (long *)&tot, *b = (long *)&st;
            for (size_t i = 0; i < sizeof(stats_t) / sizeof(long); i++) a[i] += b[i];
        }
    }
    printf("frames %ld flagged %ld (%.4f%%): nonfinite %ld cmax %ld weakcand %ld sign %ld rd %ld\n", tot.frames, tot.flagged,
           100.0 * tot.flagged / tot.frames, tot.flag_nonfinite, tot.flag_cmax, tot.flag_weakcand, tot.flag_sign, tot.flag_rd);
    printf("exact coarse lags %ld (%.3f per frame), F1~F2 ties %ld, mean |C| %.3f\n", tot.exact_lags, (double)tot.exact_lags / tot.frames,
           tot.tie12, (double)tot.csum / tot.frames);
    printf("MISMATCHES among unflagged: period %ld gain %ld\n", tot.mism_period, tot.mism_gain);
    return (tot.mism_period || tot.mism_gain) ? 1 : 0;
}
