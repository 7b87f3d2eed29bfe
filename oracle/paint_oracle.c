/*
 * paint_oracle.c — CPU restatement of Relate's chromosome-painting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * path; nothing in the product (relate_b200/, the C-ABI library, the CLI)
 * may link, import or execute it.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the
 * reference binary built from /root/reference (oracle/_ref/Relate, recipe in
 * oracle/Makefile) and requires `cmp`-identical chunk_<c>/paint/relate_<w>.bin
 * files; tests/golden/ holds reference-generated fixtures for boxes where
 * /root/reference is absent.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/include).  Arithmetic is fp64 with the reference's
 * evaluation and summation order; nothing here is copied code — the
 * reference walks iterators over CollapsedMatrix rows, this file indexes
 * flat arrays.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

/* ---- src/fast_log.hpp:6-22 -------------------------------------------- */
float ro_fast_log2(float val)
{
    int32_t x;
    memcpy(&x, &val, 4);
    const int log_2 = ((x >> 23) & 255) - 128;
    x &= ~(255 << 23);
    x += 127 << 23;
    memcpy(&val, &x, 4);
    /* separate float roundings for every operation, as an x86-64 SSE build
     * without FMA evaluates it */
    float t = (-1.0f / 3) * val;
    t = t + 2;
    t = t * val;
    t = t - 2.0f / 3;
    return t + (float)log_2;
}

float ro_fast_log(float val) { return ro_fast_log2(val) * 0.69314718f; }

/* ---- src/collapsed_matrix.hpp:228-265 (stepping-stone record writer) ---- */
/* Run-length rule: a value joins the current run when
 *   fabs(head - v) < 1e-3 * min(head, v)
 * (float subtraction, comparison in double).  Returns the number of runs. */
int ro_rle_encode(const float *v, int n, float *vals, int *lens)
{
    float head = v[0];
    int k = 0;
    vals[0] = head;
    lens[0] = 1;
    for (int j = 1; j < n; j++) {
        float x = v[j];
        float mn = (x < head) ? x : head; /* std::min(head, x) */
        if ((double)fabsf(head - x) < 1e-3 * (double)mn) {
            lens[k]++;
        } else {
            head = x;
            k++;
            vals[k] = head;
            lens[k] = 1;
        }
    }
    return k + 1;
}

static int write_record(FILE *fp, const float *v, int n, int site, float logscale,
                        float *vals, int *lens)
{
    size_t one = 1, sub = (size_t)n;
    int k = ro_rle_encode(v, n, vals, lens);
    if (fwrite(&one, sizeof(size_t), 1, fp) != 1) return -1;
    if (fwrite(&sub, sizeof(size_t), 1, fp) != 1) return -1;
    if (fwrite(&site, sizeof(int), 1, fp) != 1) return -1;
    if (fwrite(&logscale, sizeof(float), 1, fp) != 1) return -1;
    if (fwrite(&k, sizeof(int), 1, fp) != 1) return -1;
    if (fwrite(vals, sizeof(float), (size_t)k, fp) != (size_t)k) return -1;
    if (fwrite(lens, sizeof(int), (size_t)k, fp) != (size_t)k) return -1;
    return 0;
}

/* ---- src/fast_painting.cpp:18-618 PaintSteppingStones -------------------
 * hap   : L*N chars '0'/'1', SNP-major (Data::sequence)
 * r     : L doubles (Data::r, already multiplied by rho)
 * wb    : W+1 window boundaries, wb[0]=0, wb[W]=L
 * out   : alpha,beta W*N floats; ls_* W floats; site_* W ints
 * returns 0, or a negative code on inconsistent input                     */
int ro_paint_target(const char *hap, int N, int L, const double *r, double theta,
                    const int *wb, int W, int k, float *alpha, float *beta, float *ls_alpha,
                    float *ls_beta, int *site_begin, int *site_end)
{
    if (L < 2 || N < 2 || W < 1 || wb[W] != L) return -1;
    const double ntheta = 1.0 - theta;
    /* fast_painting.hpp:26-39 */
    const double lower = 1e-10, upper = 1.0 / lower;
    const double Nm1 = N - 1.0;
    const double prior_theta = theta / Nm1 - ntheta / Nm1;
    const double prior_ntheta = ntheta / Nm1;
    const double theta_ratio = theta / (1.0 - theta) - 1.0;
    const double log_ntheta = log(ntheta);
    const double log_small = log(0.01);
    const int last = L - 1;

    double *rp = (double *)malloc(sizeof(double) * (size_t)(L + 2));
    double *nor = (double *)malloc(sizeof(double) * (size_t)(L + 2));
    int *der = (int *)malloc(sizeof(int) * (size_t)(L + 2));
    double *cur = (double *)malloc(sizeof(double) * (size_t)N);
    double *prv = (double *)malloc(sizeof(double) * (size_t)N);
    if (!rp || !nor || !der || !cur || !prv) return -2;

    /* ---- site list, recombination tables, boundary sites (:41-157) ---- */
    int nbeg = 0, nend = 0, widx = 1, wend = wb[1];
    site_begin[nbeg++] = 0;
    int m = 0; /* index of the site currently being closed */
    der[0] = 0;
    rp[0] = r[0];
    int snp = 1;
    for (;;) {
        while (hap[(size_t)snp * N + k] != '1' && snp != last) {
            rp[m] += r[snp];
            snp++;
        }
        if (snp >= wend && der[m] < wend) {
            while (wend <= snp) {
                if (nend >= W || nbeg >= W) return -3;
                site_end[nend++] = snp;
                site_begin[nbeg++] = der[m];
                widx++;
                wend = wb[widx];
            }
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
        if (snp >= L) break;
    }
    /* trailing entry, x_m = r[L-1] (:132-141) */
    nor[m] = -rp[m] + log_ntheta;
    rp[m] = 1.0 - exp(-rp[m]);
    if (rp[m] > 0.99) {
        rp[m] = 0.99;
        nor[m] = log_small + log_ntheta;
    }
    rp[m + 1] = 1.0; /* :143-144 "technicality" */
    const int num_sites = m + 1;
    if (nend >= W + 1) return -3;
    site_end[nend++] = last; /* :150 */
    if (nbeg != W || nend != W) return -4;

    /* ---- forward (:201-378) ---- */
    double logscale = 0.0, sum = 0.0;
    {
        const char *row = hap;
        const char sk = row[k];
        for (int n = 0; n < N; n++) {
            double derived = (double)(sk > row[n]);
            cur[n] = derived * prior_theta + prior_ntheta;
        }
        cur[k] = 0.0;
        for (int n = 0; n < N; n++) sum += cur[n];
    }
    int wa = 0;
    while (wa < W && site_begin[wa] == 0) {
        for (int n = 0; n < N; n++) alpha[(size_t)wa * N + n] = (float)cur[n];
        ls_alpha[wa] = (float)logscale;
        wa++;
    }
    double R;
    int ir = 0; /* index into rp (it_r_prob) */
    R = rp[ir] / ((1.0 - rp[ir]) * Nm1) * sum;
    for (int i = 1; i < num_sites; i++) {
        const int s = der[i];
        const char *row = hap + (size_t)s * N;
        const char sk = row[k];
        double *t = prv; prv = cur; cur = t;
        logscale += nor[i - 1];
        for (int n = 0; n < N; n++) {
            double v = prv[n] + R;
            double derived = (double)(sk > row[n]);
            v *= derived * theta_ratio + 1.0;
            cur[n] = v;
        }
        cur[k] = 0.0;
        sum = 0.0;
        for (int n = 0; n < N; n++) sum += cur[n];
        R = sum;
        if (R < lower || R > upper) { /* :334-347 */
            double tmp = R;
            for (int n = 0; n < N; n++) cur[n] /= tmp;
            logscale += log(tmp);
            R = 1.0;
        }
        ir++;
        if (rp[ir] < 1.0) R *= rp[ir] / ((1.0 - rp[ir]) * Nm1);
        while (wa < W && site_begin[wa] == s) { /* :354-374 */
            for (int n = 0; n < N; n++) alpha[(size_t)wa * N + n] = (float)cur[n];
            ls_alpha[wa] = (float)logscale;
            wa++;
        }
    }
    if (wa != W) return -5;

    /* ---- backward (:396-582) ---- */
    const double norm = (double)log(Nm1) - num_sites * log_ntheta;
    logscale = norm;
    double bsum = 0.0;
    char sk;
    {
        const char *row = hap + (size_t)last * N;
        sk = row[k];
        for (int n = 0; n < N; n++) cur[n] = 1.0;
        for (int n = 0; n < N; n++) {
            if (sk > row[n]) bsum += theta; else bsum += ntheta;
        }
        bsum -= ntheta;
    }
    int wbk = W - 1;
    while (wbk >= 0 && site_end[wbk] == last) {
        for (int n = 0; n < N; n++) beta[(size_t)wbk * N + n] = (float)cur[n];
        ls_beta[wbk] = (float)logscale;
        wbk--;
    }
    /* ir == m here: the forward loop leaves it_r_prob on the trailing entry (:349,454) */
    R = rp[ir] / ((1.0 - rp[ir]) * Nm1) * bsum;
    int inor = m; /* it_nor_x_theta after the forward loop */
    int snp_next = last;
    for (int i = num_sites - 2; i >= 0; i--) {
        const int s = der[i];
        double *t = prv; prv = cur; cur = t;
        logscale += nor[inor];
        const double b_1mt = R / ntheta;
        const double b_t = R / theta - b_1mt;
        const char *rown = hap + (size_t)snp_next * N;
        for (int n = 0; n < N; n++) { /* :481-488, sk is the target's allele at snp_next */
            double derived = (double)(sk > rown[n]);
            double v = prv[n] + derived * b_t + b_1mt;
            v *= derived * theta_ratio + 1.0;
            cur[n] = v;
        }
        const char *row = hap + (size_t)s * N;
        sk = row[k];
        cur[k] = 0.0;
        bsum = 0.0;
        for (int n = 0; n < N; n++) { /* :495-503 */
            if (sk > row[n]) bsum += theta * cur[n]; else bsum += ntheta * cur[n];
        }
        R = bsum;
        if (R < lower || R > upper) { /* :538-551, fast_log on the float-converted sum */
            double tmp = R;
            for (int n = 0; n < N; n++) cur[n] /= tmp;
            logscale += ro_fast_log((float)tmp);
            R = 1.0;
        }
        ir--;
        if (rp[ir] < 1.0) R *= rp[ir] / ((1.0 - rp[ir]) * Nm1);
        while (wbk >= 0 && site_end[wbk] == s) { /* :559-578 */
            for (int n = 0; n < N; n++) beta[(size_t)wbk * N + n] = (float)cur[n];
            ls_beta[wbk] = (float)logscale;
            wbk--;
        }
        snp_next = s;
        inor--;
    }
    if (wbk != -1) return -6;

    free(rp); free(nor); free(der); free(cur); free(prv);
    return 0;
}

/* Number of visited sites D_k = 2 + #{s in [1,L-2] : hap[s][k]=='1'} (:52-131) */
long ro_count_sites(const char *hap, int N, int L, int k)
{
    long d = 2;
    for (int s = 1; s < L - 1; s++) d += hap[(size_t)s * N + k] == '1';
    return d;
}

/* ---- loaders: src/data.cpp:86-97,531-540; data.hpp:91-101;
 *      collapsed_matrix.hpp:215-225 ------------------------------------- */
static int file_exists(const char *p)
{
    struct stat st;
    return stat(p, &st) == 0;
}

/* ---- pipeline/Paint.cpp:17-108 -----------------------------------------
 * dir       : the -o directory
 * painting  : the --painting string ("theta,rho") or NULL when the flag is absent
 * k_begin/k_end : target range (the reference always does 0..N); the range
 *             lets bench.py time a bounded sample.
 * stats_out : [0]=N [1]=L [2]=W [3]=sum of D_k over painted targets          */
int ro_paint_chunk(const char *dir, int chunk, const char *painting, int k_begin, int k_end,
                   double *stats_out)
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

    static const char *ext[] = {"bp", "dist", "rpos", "state"};
    for (int e = 0; e < 4; e++) { /* the reference loader opens (and asserts on) all six files */
        snprintf(path, sizeof path, "%s/chunk_%d.%s", dir, chunk, ext[e]);
        if (!file_exists(path)) return -12;
    }
    snprintf(path, sizeof path, "%s/chunk_%d.hap", dir, chunk);
    fp = fopen(path, "rb");
    if (!fp) return -12;
    size_t uL, uN;
    if (fread(&uL, 8, 1, fp) != 1 || fread(&uN, 8, 1, fp) != 1) return -13;
    if ((int)uL != L || (int)uN != N) return -13;
    char *hap = (char *)malloc(uL * uN);
    if (fread(hap, 1, uL * uN, fp) != uL * uN) return -13;
    fclose(fp);
    snprintf(path, sizeof path, "%s/chunk_%d.r", dir, chunk);
    fp = fopen(path, "rb");
    if (!fp) return -12;
    unsigned rl;
    if (fread(&rl, 4, 1, fp) != 1 || (int)rl != L) return -14;
    double *r = (double *)malloc(sizeof(double) * (size_t)L);
    if (fread(r, 8, (size_t)L, fp) != (size_t)L) return -14;
    fclose(fp);

    double theta = 0.001; /* data.cpp:95 */
    if (painting) { /* Paint.cpp:38-61: both numbers go through std::stof */
        char *end;
        theta = (double)strtof(painting, &end);
        double rho = 1.0;
        if (*end == ',') rho = (double)strtof(end + 1, NULL);
        for (int l = 0; l < L; l++) r[l] *= rho;
    }

    snprintf(path, sizeof path, "%s/chunk_%d", dir, chunk);
    mkdir(path, 0700);
    snprintf(path, sizeof path, "%s/chunk_%d/paint", dir, chunk);
    mkdir(path, 0700);
    FILE **pf = (FILE **)malloc(sizeof(FILE *) * (size_t)W);
    for (int w = 0; w < W; w++) {
        snprintf(path, sizeof path, "%s/chunk_%d/paint/relate_%d.bin", dir, chunk, w);
        pf[w] = fopen(path, "wb");
        if (!pf[w]) return -15;
    }

    float *alpha = (float *)malloc(sizeof(float) * (size_t)W * N);
    float *beta = (float *)malloc(sizeof(float) * (size_t)W * N);
    float *lsa = (float *)malloc(sizeof(float) * (size_t)W);
    float *lsb = (float *)malloc(sizeof(float) * (size_t)W);
    int *sb = (int *)malloc(sizeof(int) * (size_t)W);
    int *se = (int *)malloc(sizeof(int) * (size_t)W);
    float *vals = (float *)malloc(sizeof(float) * (size_t)N);
    int *lens = (int *)malloc(sizeof(int) * (size_t)N);
    if (k_begin < 0) k_begin = 0;
    if (k_end < 0 || k_end > N) k_end = N;
    double sumD = 0;
    int rc = 0;
    for (int k = k_begin; k < k_end && rc == 0; k++) {
        rc = ro_paint_target(hap, N, L, r, theta, wb, W, k, alpha, beta, lsa, lsb, sb, se);
        if (rc) break;
        sumD += (double)ro_count_sites(hap, N, L, k);
        for (int w = 0; w < W; w++) { /* fast_painting.cpp:589-601 */
            int a = wb[w], b = wb[w + 1] - 1;
            fwrite(&a, 4, 1, pf[w]);
            fwrite(&b, 4, 1, pf[w]);
            if (write_record(pf[w], alpha + (size_t)w * N, N, sb[w], lsa[w], vals, lens)) rc = -16;
            if (write_record(pf[w], beta + (size_t)w * N, N, se[w], lsb[w], vals, lens)) rc = -16;
        }
    }
    for (int w = 0; w < W; w++) fclose(pf[w]);
    if (stats_out) {
        stats_out[0] = N; stats_out[1] = L; stats_out[2] = W; stats_out[3] = sumD;
    }
    free(alpha); free(beta); free(lsa); free(lsb); free(sb); free(se); free(vals); free(lens);
    free(pf); free(hap); free(r); free(wb);
    return rc;
}
