/*
 * synth.c — deterministic synthetic haplotype generator (block-Kingman).
 *
 * Used by bench.py and the tests to make inputs of the shapes BASELINE.json names
 * (there is no msprime here and no network).  Blocks of `block` consecutive SNPs
 * share one Kingman coalescent tree on N leaves; each SNP is one mutation on a
 * branch drawn proportionally to branch length, carriers = the leaves below it.
 *
 * The ranked topology is drawn as a random Cartesian tree: put the leaves in a
 * uniformly random order on a line and remove the N-1 gaps between neighbours in
 * uniformly random order; run forwards this is the Yule process, whose ranked
 * shapes are Kingman's.  An internal node is then a contiguous interval of the
 * line, so carriers are a slice of the leaf permutation.
 *
 * Output layout is the reference's Data::sequence (SNP-major chars '0'/'1',
 * include/src/data.cpp:531-540) so the bytes can be written straight into a
 * chunk_<c>.hap file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t splitmix(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t next(rng_t *r)
{
    uint64_t *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static double unif(rng_t *r) { return ((next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static uint32_t below(rng_t *r, uint32_t n) { return (uint32_t)(((next(r) >> 32) * (uint64_t)n) >> 32); }

/* hap: L*N chars; bp: L ints (strictly increasing positions). returns 0 */
int synth_block_kingman(int N, int L, int block, uint64_t seed, char *hap, int *bp)
{
    rng_t rg;
    uint64_t sm = seed * 0x2545F4914F6CDD1DULL + 12345;
    for (int i = 0; i < 4; i++) rg.s[i] = splitmix(&sm);

    const int G = N - 1; /* gaps == internal nodes */
    int *perm = malloc(sizeof(int) * N);
    int *rank = malloc(sizeof(int) * G);   /* removal order of gap g (0 = first merge) */
    int *lo = malloc(sizeof(int) * G), *hi = malloc(sizeof(int) * G); /* leaf interval [lo,hi] */
    int *par = malloc(sizeof(int) * G);    /* parent gap, -1 for root */
    int *stack = malloc(sizeof(int) * (G + 1));
    double *tm = malloc(sizeof(double) * G);   /* time of j-th merge */
    double *cum = malloc(sizeof(double) * (2 * N)); /* cumulative branch lengths: leaves then gaps */
    if (!perm || !rank || !lo || !hi || !par || !stack || !tm || !cum) return -1;

    int pos = 0;
    for (int s0 = 0; s0 < L; s0 += block) {
        for (int i = 0; i < N; i++) perm[i] = i;
        for (int i = N - 1; i > 0; i--) { int j = below(&rg, i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
        for (int i = 0; i < G; i++) rank[i] = i;
        for (int i = G - 1; i > 0; i--) { int j = below(&rg, i + 1); int t = rank[i]; rank[i] = rank[j]; rank[j] = t; }
        double t = 0;
        for (int j = 0; j < G; j++) { /* k = N-j lineages */
            double k = N - j;
            t += -log(unif(&rg)) / (k * (k - 1) / 2.0);
            tm[j] = t;
        }
        /* nearest gap with larger rank on each side (Cartesian tree by rank, max at root) */
        int sp = 0;
        for (int g = 0; g < G; g++) {
            while (sp > 0 && rank[stack[sp - 1]] < rank[g]) sp--;
            int left = sp > 0 ? stack[sp - 1] : -1;
            lo[g] = left + 1;      /* leaves left+1 .. */
            par[g] = left;         /* provisional */
            stack[sp++] = g;
        }
        sp = 0;
        for (int g = G - 1; g >= 0; g--) {
            while (sp > 0 && rank[stack[sp - 1]] < rank[g]) sp--;
            int right = sp > 0 ? stack[sp - 1] : -1;
            hi[g] = right >= 0 ? right : G; /* .. up to leaf index `right` (gap r sits between leaf r and r+1) */
            int left = par[g];
            if (left < 0) par[g] = right;
            else if (right >= 0 && rank[right] < rank[left]) par[g] = right;
            stack[sp++] = g;
        }
        double acc = 0;
        for (int p = 0; p < N; p++) { /* leaf branches */
            int a = p - 1, b = p, pg;
            if (a < 0) pg = b; else if (b >= G) pg = a; else pg = rank[a] < rank[b] ? a : b;
            acc += tm[rank[pg]];
            cum[p] = acc;
        }
        for (int g = 0; g < G; g++) { /* internal branches; root has length 0 */
            if (par[g] >= 0) acc += tm[rank[par[g]]] - tm[rank[g]];
            cum[N + g] = acc;
        }
        int s1 = s0 + block < L ? s0 + block : L;
        for (int s = s0; s < s1; s++) {
            double u = unif(&rg) * acc;
            int a = 0, b = N + G - 1;
            while (a < b) { int mid = (a + b) >> 1; if (cum[mid] > u) b = mid; else a = mid + 1; }
            char *row = hap + (size_t)s * N;
            memset(row, '0', N);
            if (a < N) row[perm[a]] = '1';
            else { int g = a - N; for (int p = lo[g]; p <= hi[g]; p++) row[perm[p]] = '1'; }
            pos += 1 + (int)below(&rg, 199);
            bp[s] = pos;
        }
    }
    free(perm); free(rank); free(lo); free(hi); free(par); free(stack); free(tm); free(cum);
    return 0;
}
