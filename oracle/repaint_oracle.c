/*
 * repaint_oracle.c — CPU restatement of the consumer side of the painting path:
 * FastPainting::RePaintSection (src/fast_painting.cpp:620-1092) and
 * DistanceMeasure::GetTopologyWithRepaint / GetMatrix (src/anc_builder.cpp:48-207).
 *
 * TEST INFRASTRUCTURE ONLY (see paint_oracle.c).  Parity status: PINNED — tests/test_repaint_cpu.py requires
 * ro_window_distances() to reproduce, byte for byte, what oracle/_ref/dlens (our driver around the UNMODIFIED
 * reference's GetMatrix) writes for the same paint files.
 *
 * fp64 state, the reference's evaluation order, float/double mixing as written there.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

float ro_fast_log(float val); /* paint_oracle.c */

/* ---- RePaintSection -------------------------------------------------------------------------------
 * alpha_begin/beta_end: N floats (decoded stepping stones of window w for target k); bS/eS their sites.
 * top: out, (m+1)*N floats (caller sizes it for eS-bS+2 rows); logscales: out, m+1 floats; returns m+1. */
int ro_repaint_section(const char *hap, int N, int L, const double *r, double theta, const float *alpha_begin,
                       const float *beta_end, int bS, int eS, float ls_alpha, float ls_beta, int k, float *top,
                       float *logscales)
{
    (void)L;
    const double ntheta = 1.0 - theta;
    const double lower = 1e-10, upper = 1.0 / lower; /* fast_painting.hpp:26-39 */
    const double Nm1 = N - 1.0;
    const double theta_ratio = theta / (1.0 - theta) - 1.0;
    const double log_ntheta = log(ntheta);
    const double log_small = log(0.01);
    const int max_snps = eS - bS + 2;
    double *rp = (double *)malloc(sizeof(double) * (size_t)(max_snps + 1));
    double *nor = (double *)malloc(sizeof(double) * (size_t)(max_snps + 1));
    int *der = (int *)malloc(sizeof(int) * (size_t)(max_snps + 1));

    /* site list and recombination tables (:640-716) */
    int m = 0;
    der[0] = bS;
    rp[0] = r[bS];
    int snp = bS + 1;
    for (;;) {
        while (hap[(size_t)snp * N + k] != '1' && snp != eS) {
            rp[m] += r[snp];
            snp++;
        }
        nor[m] = -rp[m] + log_ntheta;
        rp[m] = 1.0 - exp(-rp[m]);
        if (rp[m] > 0.99) {
            rp[m] = 0.99;
            nor[m] = log_small + log_ntheta;
        }
        m++;
        der[m] = snp;
        rp[m] = r[snp];
        snp++;
        if (snp > eS) break;
    }
    nor[m] = -rp[m] + log_ntheta;
    rp[m] = 1.0 - exp(-rp[m]);
    if (rp[m] > 0.99) {
        rp[m] = 0.99;
        nor[m] = log_small + log_ntheta;
    }
    rp[m + 1] = 1.0;
    const int D = m + 1;

    double *alpha = (double *)malloc(sizeof(double) * (size_t)D * N);
    double *bcur = (double *)malloc(sizeof(double) * (size_t)N);
    double *bnext = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < D; i++) logscales[i] = 0.0f;

    /* forward (:752-885) */
    logscales[0] = ls_alpha;
    double sum = 0.0;
    for (int n = 0; n < N; n++) alpha[n] = (double)alpha_begin[n];
    alpha[k] = 0.0;
    for (int n = 0; n < N; n++) sum += alpha[n];
    int ir = 0;
    double R = rp[ir] / ((1.0 - rp[ir]) * Nm1) * sum;
    double prev_ls = (double)logscales[0];
    for (int i = 1; i < D; i++) {
        const int s = der[i];
        const char *row = hap + (size_t)s * N;
        const char sk = row[k];
        prev_ls += nor[i - 1];
        logscales[i] = (float)prev_ls;
        double *a = alpha + (size_t)i * N;
        const double *ap = alpha + (size_t)(i - 1) * N;
        for (int n = 0; n < N; n++) {
            double v = ap[n] + R;
            double derived = (double)(sk > row[n]);
            v *= derived * theta_ratio + 1.0;
            a[n] = v;
        }
        a[k] = 0.0;
        sum = 0.0;
        for (int n = 0; n < N; n++) sum += a[n];
        R = sum;
        if (R < lower || R > upper) {
            double tmp = R;
            for (int n = 0; n < N; n++) a[n] /= tmp;
            prev_ls += log(tmp);
            logscales[i] += log(tmp); /* float += double */
            R = 1.0;
        }
        ir++;
        if (rp[ir] < 1.0) R *= rp[ir] / ((1.0 - rp[ir]) * Nm1);
    }

    /* backward (:887-1073) */
    logscales[m] += ls_beta;
    double bsum = 0.0;
    char sk;
    {
        const char *row = hap + (size_t)eS * N;
        sk = row[k];
        for (int n = 0; n < N; n++) bcur[n] = (double)beta_end[n];
        bcur[k] = 0.0;
        for (int n = 0; n < N; n++) {
            if (sk > row[n]) bsum += theta * bcur[n]; else bsum += ntheta * bcur[n];
        }
        const double *a = alpha + (size_t)m * N;
        float *t = top + (size_t)m * N;
        for (int n = 0; n < N; n++) t[n] = (float)(a[n] * bcur[n]);
    }
    R = rp[ir] / ((1.0 - rp[ir]) * Nm1) * bsum; /* ir == m */
    int snp_next = eS;
    int inor = m;
    prev_ls = (double)ls_beta;
    for (int i = m - 1; i >= 0; i--) {
        const int s = der[i];
        double *tmpp = bnext; bnext = bcur; bcur = tmpp;
        prev_ls += nor[inor];
        logscales[i] += prev_ls; /* float += double */
        const double b_1mt = R / ntheta;
        const double b_t = R / theta - b_1mt;
        const char *rown = hap + (size_t)snp_next * N;
        for (int n = 0; n < N; n++) {
            double derived = (double)(sk > rown[n]);
            double v = bnext[n] + derived * b_t + b_1mt;
            v *= derived * theta_ratio + 1.0;
            bcur[n] = v;
        }
        const char *row = hap + (size_t)s * N;
        sk = row[k];
        bcur[k] = 0.0;
        bsum = 0.0;
        for (int n = 0; n < N; n++) {
            if (sk > row[n]) bsum += theta * bcur[n]; else bsum += ntheta * bcur[n];
        }
        R = bsum;
        {
            const double *a = alpha + (size_t)i * N;
            float *t = top + (size_t)i * N;
            for (int n = 0; n < N; n++) t[n] = (float)(a[n] * bcur[n]); /* taken BEFORE the rescale (:1033-1045) */
        }
        if (R < lower || R > upper) {
            double tmp = R;
            for (int n = 0; n < N; n++) bcur[n] /= tmp;
            prev_ls += log(tmp);
            logscales[i] += log(tmp);
            R = 1.0;
        }
        ir--;
        if (rp[ir] < 1.0) R *= rp[ir] / ((1.0 - rp[ir]) * Nm1);
        snp_next = s;
        inor--;
    }
    free(rp); free(nor); free(der); free(alpha); free(bcur); free(bnext);
    return D;
}

/* ---- GetMatrix (src/anc_builder.cpp:108-207) for one target row n ------------------------------------
 * top_n: n's posterior rows, ls_n its log-scales, v = v_snp_prev[n]; rpos_prev/next as kept by the caller.
 * out: N floats (row n of the distance matrix). */
void ro_matrix_row(const char *hap, int N, int L, const double *rpos, int snp, int n, const float *top_n,
                   const float *ls_n, int v, double rpos_prev, double *rpos_next_io, float *out)
{
    const float scale = -1.0f;
    float min = INFINITY;
    if (hap[(size_t)snp * N + n] == '1' || snp == 0 || snp == L - 1) {
        const float *t = top_n + (size_t)v * N;
        const float lsp = ls_n[v];
        for (int j = 0; j < N; j++) {
            out[j] = (ro_fast_log(t[j]) + lsp) * scale;
            if (out[j] < min) min = out[j];
        }
        out[n] = 0.0f;
    } else {
        if (*rpos_next_io <= rpos_prev) {
            for (int l = snp; l < L; l++) {
                if (hap[(size_t)l * N + n] == '1' || l == L - 1) {
                    *rpos_next_io = rpos[l];
                    break;
                }
            }
        }
        const double rpos_next = *rpos_next_io;
        double wl, wr;
        if (rpos_prev == rpos_next) {
            wl = 0.5; wr = 0.5;
        } else {
            const double den = rpos_next - rpos_prev;
            wl = (rpos_next - rpos[snp]) / den;
            wr = (rpos[snp] - rpos_prev) / den;
        }
        const float *tp = top_n + (size_t)v * N, *tn = top_n + (size_t)(v + 1) * N;
        const float lsp = ls_n[v], lsn = ls_n[v + 1];
        const float e_pn = expf(lsp - lsn), e_np = expf(lsn - lsp);
        for (int j = 0; j < N; j++) {
            float arg;
            float m;
            if (lsp <= lsn) {
                arg = (float)(wl * tp[j] * e_pn + wr * tn[j]);
                m = (ro_fast_log(arg) + lsn) * scale;
            } else {
                arg = (float)(wl * tp[j] + wr * tn[j] * e_np);
                m = (ro_fast_log(arg) + lsp) * scale;
            }
            out[j] = m;
            if (m < min) min = m;
        }
        out[n] = 0.0f;
    }
    for (int j = 0; j < N; j++)
        if (j != n) out[j] -= min;
}

/* ---- driver: same walk and same output format as oracle/dlens.cpp ----------------------------------- */
static int read_record(FILE *fp, int N, float *v, int *site, float *ls)
{
    size_t one, sub;
    int k;
    if (fread(&one, 8, 1, fp) != 1 || fread(&sub, 8, 1, fp) != 1) return -1;
    if (one != 1 || (int)sub != N) return -1;
    if (fread(site, 4, 1, fp) != 1 || fread(ls, 4, 1, fp) != 1 || fread(&k, 4, 1, fp) != 1) return -1;
    float *vals = (float *)malloc(sizeof(float) * (size_t)k);
    int *lens = (int *)malloc(sizeof(int) * (size_t)k);
    if (fread(vals, 4, (size_t)k, fp) != (size_t)k || fread(lens, 4, (size_t)k, fp) != (size_t)k) return -1;
    int i = 0;
    for (int j = 0; j < k; j++)
        for (int t = 0; t < lens[j] && i < N; t++) v[i++] = vals[j];
    free(vals); free(lens);
    return i == N ? 0 : -1;
}

int ro_window_distances(const char *dir, int chunk, int section, int stride, const char *painting, const char *out_path)
{
    char path[4096];
    int N, L, nb;
    snprintf(path, sizeof path, "%s/parameters_c%d.bin", dir, chunk);
    FILE *fp = fopen(path, "rb");
    if (!fp) return -10;
    if (fread(&N, 4, 1, fp) != 1 || fread(&L, 4, 1, fp) != 1 || fread(&nb, 4, 1, fp) != 1) return -11;
    int *wb = (int *)malloc(sizeof(int) * (size_t)nb);
    if (fread(wb, 4, (size_t)nb, fp) != (size_t)nb) return -11;
    fclose(fp);
    const int W = nb - 1;
    if (section < 0 || section >= W) return -12;
    snprintf(path, sizeof path, "%s/chunk_%d.hap", dir, chunk);
    fp = fopen(path, "rb");
    if (!fp) return -13;
    size_t uL, uN;
    if (fread(&uL, 8, 1, fp) != 1 || fread(&uN, 8, 1, fp) != 1) return -13;
    char *hap = (char *)malloc(uL * uN);
    if (fread(hap, 1, uL * uN, fp) != uL * uN) return -13;
    fclose(fp);
    unsigned n32;
    snprintf(path, sizeof path, "%s/chunk_%d.r", dir, chunk);
    fp = fopen(path, "rb");
    if (!fp || fread(&n32, 4, 1, fp) != 1) return -14;
    double *r = (double *)malloc(sizeof(double) * (size_t)L);
    if (fread(r, 8, (size_t)L, fp) != (size_t)L) return -14;
    fclose(fp);
    snprintf(path, sizeof path, "%s/chunk_%d.rpos", dir, chunk);
    fp = fopen(path, "rb");
    if (!fp || fread(&n32, 4, 1, fp) != 1) return -15;
    double *rpos = (double *)malloc(sizeof(double) * (size_t)(L + 1));
    if (fread(rpos, 8, (size_t)(L + 1), fp) != (size_t)(L + 1)) return -15;
    fclose(fp);
    double theta = 0.001;
    if (painting && strcmp(painting, "-") != 0) {
        char *end;
        theta = (double)strtof(painting, &end);
        double rho = (*end == ',') ? (double)strtof(end + 1, NULL) : 1.0;
        for (int l = 0; l < L; l++) r[l] *= rho;
    }
    const int start = wb[section];
    const int end = (section < W - 1) ? wb[section + 1] - 1 : L - 1;

    /* GetTopologyWithRepaint (anc_builder.cpp:48-106) at snp = start */
    snprintf(path, sizeof path, "%s/chunk_%d/paint/relate_%d.bin", dir, chunk, section);
    fp = fopen(path, "rb");
    if (!fp) return -16;
    float **top = (float **)calloc((size_t)N, sizeof(float *));
    float **ls = (float **)calloc((size_t)N, sizeof(float *));
    float *ab = (float *)malloc(sizeof(float) * (size_t)N), *be = (float *)malloc(sizeof(float) * (size_t)N);
    for (int n = 0; n < N; n++) {
        int a, b, bS, eS;
        float lsa, lsb;
        if (fread(&a, 4, 1, fp) != 1 || fread(&b, 4, 1, fp) != 1) return -17;
        if (read_record(fp, N, ab, &bS, &lsa) || read_record(fp, N, be, &eS, &lsb)) return -17;
        if (bS > a || eS < b) return -18;
        top[n] = (float *)malloc(sizeof(float) * (size_t)(eS - bS + 2) * N);
        ls[n] = (float *)malloc(sizeof(float) * (size_t)(eS - bS + 2));
        ro_repaint_section(hap, N, L, r, theta, ab, be, bS, eS, lsa, lsb, n, top[n], ls[n]);
    }
    fclose(fp);
    int *v = (int *)calloc((size_t)N, sizeof(int));
    double *rp_prev = (double *)malloc(sizeof(double) * (size_t)N), *rp_next = (double *)malloc(sizeof(double) * (size_t)N);
    {
        const int snp = start;
        if (snp > 0)
            for (int n = 0; n < N; n++) v[n] += hap[(size_t)snp * N + n] == '1'; /* tsnp runs from snp down to start */
        for (int n = 0; n < N; n++) {
            int t = snp;
            while (hap[(size_t)t * N + n] != '1' && t > 0) t--;
            rp_prev[n] = rpos[t];
            rp_next[n] = rp_prev[n];
        }
    }
    FILE *out = fopen(out_path, "wb");
    if (!out) return -19;
    int count = 0;
    fwrite(&N, 4, 1, out);
    fwrite(&count, 4, 1, out);
    float *row = (float *)malloc(sizeof(float) * (size_t)N);
    for (int snp = start; snp <= end; snp++) {
        if (snp > start) { /* anc_builder.cpp:487-495 */
            for (int n = 0; n < N; n++) {
                if (hap[(size_t)snp * N + n] == '1') {
                    v[n]++;
                    rp_prev[n] = rpos[snp];
                }
            }
        }
        if (snp == start || (snp - start) % stride == 0 || snp == end) {
            fwrite(&snp, 4, 1, out);
            for (int n = 0; n < N; n++) {
                ro_matrix_row(hap, N, L, rpos, snp, n, top[n], ls[n], v[n], rp_prev[n], &rp_next[n], row);
                fwrite(row, 4, (size_t)N, out);
            }
            count++;
        }
    }
    fseek(out, 4, SEEK_SET);
    fwrite(&count, 4, 1, out);
    fclose(out);
    for (int n = 0; n < N; n++) { free(top[n]); free(ls[n]); }
    free(top); free(ls); free(ab); free(be); free(v); free(rp_prev); free(rp_next); free(row);
    free(hap); free(r); free(rpos); free(wb);
    return 0;
}
