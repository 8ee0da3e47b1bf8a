/* Minimal C client of libafb200.so: the drop-in boundary is a plain C ABI (include/anchored_fusion.h).
 *
 *   gcc -std=c99 -Iinclude examples/anchor_host_batch.c -o anchor_host_batch \
 *       -Lanchored_fusion_b200 -lafb200 -Wl,-rpath,$PWD/anchored_fusion_b200
 *   ./anchor_host_batch            # without a B200 it stops at af_index_upload with AF_ERR_CUDA
 *
 * It builds an index for a toy anchor (what `bwa index` does in the reference, Anchored_Fusion.py:172),
 * packs two read pairs, streams them through the GPU path (what `bwa mem -M | samtools view -F 772`
 * does, Anchored_Fusion.py:182,194) and prints the anchored records. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "anchored_fusion.h"

static const char *ANCHOR =
    "ATGGTGGACCCGGTGGGCTTCGCGGAGGCGTGGAAGGCGCAGTTCCCGGACTCAGAGCCCCCGCGCATGGAGCTGCGCTCAGTGGGCGACATCGAGCAGGAGCTGGAGCGCTGCAAGGCCTCCATTCGGCGCCTGGAGCAGGAGGTGAACCAGGAGCGCTTCCGCATGATCTACCTGCAGACGTTGCTGGCCAAGGAAAAGAAGAGCTATGACCGGCAGCGATGGGGCTTCCGGCGCGCGGCGCAGGCCCCCGACGGCGCCTCCGAGCCCCGAGCGTCCGCGTCGCGCCCGCAGCCAGCGCCCGCCGACGGAGCCGACCCGCCGCCCGCCGAGGAGCCCGAGGCCCGGCCCGACGGCGAGGGTTCTCCGGGTAAGGCCAGGCCCGGGACCGCCCGCAGGCCCGGGGCAGCCGCGTCGGGGGAACGGGACGACCGGGGACCCCCCGCCAGCGTGGCGGCGCTCAGGTCCAACTTCGAGCGGATCCGCAAGGGCCATGGCCAGCCCGGGGCGGACGCCGAGAAGCCCTTCTACGTGAACGTCGAGTTTCACCACGAGCGCGGCCTGGTGAAGGTCAACGACAAAGAGGTGTCGGACCGCATCAGCTCCCTGGGCAGCCAGGCCATGCAGATGGAGCGCAAAAAGTCCCAGCACGGCGCGGGCTCGAGCGTGGGGGATGCATCCAGGCCCCCTTACCGGGGACGCTCCTCGGAGAGCAGCTGCGGCGTCGACGGCGACTACGAGGACGCCGAGTTGAACCCCCGCTTCCTGAAGGACAACCTGATCGACGCCAATGGCGGTAGCAGGCCCCCTTGGCCGCCCCTGGAGTACCAGCCCTACCAGAGCATCTACGTCGGGGGCATGATGGAAGGGGAGGGCAAGGGCCCGCTCCTGCGCAGCCAGAGCACCTCTGAGCAGGAGAAGCGCCTTACCTGGCCCCGCAGGTCCTACTCCCCCCGGAGTTTTGAGGATTGCGGAGGCGGCTATACCCCGGACTGCAGCTCCAATGAGAACCTCACCTCCAGCGAGGAGGACTTCTCCTCTGGCCAGTCCAGCCGCGTGTCCCCAAGCCCCACCACCTACCGCATGTTCCGGGACAAAAGCCGCTCTCCCTCGCAGAACTCGCAACAGTCCTTCGACAGCAGCAGTCCCCCCACGCCGCAGTGCCATAAGCGGCACCGGCACTGCCCGGTTGTCGTGTCCGAGGCCACCATCGTGGGCGTCCGCAAGACCGGGCAGATCTGGCCCAACGATGGCGAGGGCGCCTTCCATGGAGACGCAG";

static void die(const char *what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, af_last_error());
    exit(rc == AF_ERR_CUDA ? 3 : 1);
}

int main(void) {
    const int L = 100;
    af_index_t *idx = NULL;
    af_index_info_t info;
    int rc = af_index_build(ANCHOR, (int64_t)strlen(ANCHOR), NULL, 0, &idx);
    if (rc) die("af_index_build", rc);
    af_index_info(idx, &info);
    printf("ABI %d, anchor %d bp, k'=%d, %d filter buckets, pad byte 0x%02X\n", af_abi_version(), info.anchor_len,
           info.kp, info.n_buckets, info.pad_byte);

    /* pair 0: mate 1 = anchor[200,300), mate 2 = reverse complement of anchor[400,500); pair 1: unrelated */
    char reads1[2][101], reads2[2][101];
    memcpy(reads1[0], ANCHOR + 200, (size_t)L);
    for (int i = 0; i < L; i++) {
        char c = ANCHOR[400 + L - 1 - i];
        reads2[0][i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        reads1[1][i] = "ACGT"[(i * 7 + 3) % 4];
        reads2[1][i] = "ACGT"[(i * 5 + 1) % 4];
    }
    char cat1[200], cat2[200];
    int64_t off[3] = {0, L, 2 * L};
    memcpy(cat1, reads1[0], (size_t)L); memcpy(cat1 + L, reads1[1], (size_t)L);
    memcpy(cat2, reads2[0], (size_t)L); memcpy(cat2 + L, reads2[1], (size_t)L);

    af_layout_t lay;
    if ((rc = af_layout(L, 2, &lay))) die("af_layout", rc);
    void *packed = af_host_alloc((size_t)lay.packed_bytes);          /* pinned when a GPU is there */
    if (!packed) packed = calloc(1, (size_t)lay.packed_bytes);
    uint16_t lens[4];
    uint32_t nids[4], nmask[4 * AF_NMASK_WORDS];
    int64_t n_n = 0;
    int32_t ulen = 0;
    rc = af_pack_pairs(cat1, off, cat2, off, 2, L, info.pad_byte, packed, lens, nids, nmask, 4, &n_n, &ulen);
    if (rc) die("af_pack_pairs", rc);

    af_dev_index_t *didx = NULL;
    if ((rc = af_index_upload(idx, 0, &didx))) die("af_index_upload", rc);
    af_pipeline_t *pipe = NULL;
    if ((rc = af_pipeline_create(didx, 1024, L, 2, &pipe))) die("af_pipeline_create", rc);
    af_batch_t batch = {packed, 2, L, ulen, lens, NULL, NULL, 0};
    af_hit_t hits[4];
    int64_t n_hits = 0, n_flagged = 0;
    if ((rc = af_pipeline_run(pipe, &batch, hits, 4, &n_hits, &n_flagged))) die("af_pipeline_run", rc);
    for (int64_t i = 0; i < n_hits; i++)
        printf("read %u (pair %u mate %u): POS %d  %uS%uM%uS  strand %u score %u\n", hits[i].read_id, hits[i].read_id >> 1,
               (hits[i].read_id & 1) + 1, hits[i].pos, hits[i].clip_l, hits[i].m_len, hits[i].clip_r,
               hits[i].score_strand & 1u, hits[i].score_strand >> 1);
    af_pipeline_free(pipe);
    af_dev_index_free(didx);
    af_index_free(idx);
    return n_hits == 2 ? 0 : 2;
}
