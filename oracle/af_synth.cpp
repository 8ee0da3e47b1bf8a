// af_synth.cpp -- the synthetic read-pair generator for the CPU arm (bench.py --impl reference, cpu_baseline).
//
// TEST / MEASUREMENT INFRASTRUCTURE, not product code.  It exists so that the reference arm of bench.py
// loads no product library: the generator is a pure function of (seed, pair index) defined inline in
// anchored_fusion_b200/csrc/af_common.h (af_make_frag / af_read_base / af_ref_base), shared with the
// device generator k_synth_pairs, so both arms see the same pairs.  Stands in for the reference's
// utils/simulate_reads.py:20 (wgsim -d 200 -1 101 -2 101; wgsim is absent and the configs ask for 2x150).
#include <cstdarg>
#include <cstdint>
#include "../anchored_fusion_b200/csrc/af_common.h"

void af_set_error(const char *, ...) {}

extern "C" int afo_synth_anchor(const af_synth_t *s, char *ascii_out) {
    if (!s || !ascii_out || s->anchor_len <= 0) return -1;
    for (int32_t i = 0; i < s->anchor_len; i++) ascii_out[i] = "ACGT"[af_ref_base(s->seed, s->anchor_start + i)];
    return 0;
}

// codes 0..4, reads interleaved: row 2p = mate 1 of pair first_pair+p, row 2p+1 = mate 2; `stride` bytes per row
extern "C" int afo_synth_reads(const af_synth_t *s, int64_t first_pair, int64_t n_pairs, uint8_t *reads, int64_t stride,
                               int threads) {
    if (!s || !reads || n_pairs < 0 || stride < s->read_len) return -1;
    const int L = s->read_len;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t p = 0; p < n_pairs; p++) {
        af_frag f = af_make_frag(*s, first_pair + p);
        uint8_t *r1 = reads + (2 * p) * stride, *r2 = r1 + stride;
        for (int i = 0; i < L; i++) {
            r1[i] = (uint8_t)af_read_base(*s, f, 0, i);
            r2[i] = (uint8_t)af_read_base(*s, f, 1, i);
        }
    }
    return 0;
}
