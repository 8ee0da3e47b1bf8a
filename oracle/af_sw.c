/*
 * af_sw.c -- exhaustive affine-gap alignment of one read against the whole anchor.  TEST INFRASTRUCTURE ONLY
 * (tests/test_oracle_vs_sw.py); nothing under anchored_fusion_b200/ may use it.
 *
 * Why it exists.  The anchoring spec (af_oracle.c) restates bwa-mem for the indel-free case: exact seed, UNGAPPED
 * extension, and bwa's end-clipping rule applied greedily per side.  bwa is not available to confirm that arithmetic,
 * so this file states what bwa-mem's scoring model optimises from first principles, with no seeds, no diagonals and
 * no X-drop: over ALL local alignments of a read substring to an anchor substring with affine gaps
 * (match +A, mismatch -B, a gap of length g costs O + g*E; bwa-mem defaults 1, 4, 6, 1) maximise
 *     score + clip5 * [the alignment starts at the read's first base] + clip3 * [it ends at the read's last base]
 * i.e. clipping an end costs the clip penalty (-L 5,5).  The oracle's record of a read corresponds to the value
 *     score + clip5 * [clip_l == 0] + clip3 * [clip_r == 0]
 * of ONE ungapped alignment, so this objective can never be smaller than the oracle's, and for reads without indels
 * it must be equal unless a gap happens to pay (reported by the test).  O(L * G) per orientation, Gotoh's recurrences.
 */
#include <stdint.h>
#include <stdlib.h>

#define NEG (-(1 << 28))

/* q: oriented read codes (0..3, 4 = N), a: anchor codes.  Returns the objective; *score_out = the alignment's own score,
 * qb_out / qe_out = read interval [qb, qe), ae_out = anchor position (0-based, exclusive) where it ends. */
int afo_sw_clip_objective(const uint8_t *q, int32_t L, const uint8_t *a, int32_t G, int A, int B, int O, int E,
                          int clip5, int clip3, int32_t *score_out, int32_t *qb_out, int32_t *qe_out, int32_t *ae_out) {
    if (L <= 0 || G <= 0) return 0;
    /* H: alignment ending with q[i-1] aligned to a[j-1]; D: ending with a gap that consumes read bases (deletion from the
     * anchor's view is F).  start[i][j] is tracked to recover the read start (for reporting only). */
    int32_t *H0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(G + 1) * 6);
    int32_t *H1 = H0 + (G + 1), *F0 = H1 + (G + 1), *F1 = F0 + (G + 1), *S0 = F1 + (G + 1), *S1 = S0 + (G + 1);
    /* S: read start (0-based) of the alignment in H; SF: same for F rows kept implicitly by copying */
    int32_t *SF0 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(G + 1) * 2), *SF1 = SF0 + (G + 1);
    for (int32_t j = 0; j <= G; j++) { H0[j] = NEG; F0[j] = NEG; S0[j] = 0; SF0[j] = 0; }
    int best = NEG, best_sc = 0, bqb = 0, bqe = 0, bae = 0;
    for (int32_t i = 1; i <= L; i++) {
        const int start_val = (i == 1) ? clip5 : 0;          /* a fresh alignment beginning at read base i-1 */
        int32_t Erow = NEG, SE = 0;                          /* gap consuming anchor bases, along the row */
        H1[0] = NEG; F1[0] = NEG; S1[0] = 0; SF1[0] = 0;
        for (int32_t j = 1; j <= G; j++) {
            const int s = (q[i - 1] < 4 && q[i - 1] == a[j - 1]) ? A : -B;
            /* diagonal: extend, or start here */
            int h = start_val, st = i - 1;
            if (H0[j - 1] > h) { h = H0[j - 1]; st = S0[j - 1]; }
            if (F0[j - 1] > h) { h = F0[j - 1]; st = SF0[j - 1]; }
            /* (the row gap E of the previous row/col feeds through H of the cell it ended in: handled below) */
            h += s;
            /* gap consuming read bases (vertical): open from H0[j] or extend F0[j] */
            int f = H0[j] - O - E, sf = S0[j];
            if (F0[j] - E > f) { f = F0[j] - E; sf = SF0[j]; }
            /* gap consuming anchor bases (horizontal): open from H1[j-1] or extend */
            int e = H1[j - 1] - O - E, se = S1[j - 1];
            if (Erow - E > e) { e = Erow - E; se = SE; }
            Erow = e; SE = se;
            /* a cell's H may also END a horizontal gap (alignment ends with read base i-1 matched earlier): fold e into h's
             * successor candidates by letting H1[j] carry max(h, e) only for extension purposes */
            int hh = h, sh = st;
            if (e > hh) { hh = e; sh = se; }
            H1[j] = hh; S1[j] = sh;
            F1[j] = f; SF1[j] = sf;
            /* objective: only alignments ending in a match/mismatch column are candidates (h), not ones ending in a gap */
            const int obj = h + (i == L ? clip3 : 0);
            if (obj > best) {
                best = obj; bqb = st; bqe = i; bae = j;
                best_sc = h - (st == 0 ? clip5 : 0);
            }
        }
        int32_t *t;
        t = H0; H0 = H1; H1 = t; t = F0; F0 = F1; F1 = t; t = S0; S0 = S1; S1 = t; t = SF0; SF0 = SF1; SF1 = t;
    }
    if (score_out) *score_out = best_sc;
    if (qb_out) *qb_out = bqb;
    if (qe_out) *qe_out = bqe;
    if (ae_out) *ae_out = bae;
    /* H0 may have been swapped: free the lower of the two base pointers */
    {
        int32_t *base = H0 < H1 ? H0 : H1;
        if (F0 < base) base = F0;
        if (F1 < base) base = F1;
        if (S0 < base) base = S0;
        if (S1 < base) base = S1;
        free(base);
        free(SF0 < SF1 ? SF0 : SF1);
    }
    return best;
}
