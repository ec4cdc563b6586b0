/* TEST INFRASTRUCTURE ONLY: cg_mask_lc_bits (the bit-parallel form the device runs) and cg_mask_lc_lean (list-free, base by base) against cg_mask_lc (find_STR + add_rep with the
 * full repeat list, itself pinned to the reference's str_finder.c by the golden vectors) on random and repeat-rich reads. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../crumble_b200/csrc/cg_core.h"

static uint64_t rs = 88172645463325252ULL;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char **argv) {
    const long N = argc > 1 ? atol(argv[1]) : 200000;
    static const uint8_t nt[4] = { 1, 2, 4, 8 };
    long bad = 0, nonempty = 0;
    std::vector<uint8_t> seq4(4096), win(600);
    CgRepList *L = new CgRepList();
    for (long it = 0; it < N; it++) {
        const int mode = rnd() % 4;
        int l = mode == 3 ? 300 + rnd() % 1200 : (rnd() % 8 == 0 ? 1 + rnd() % 40 : 100 + rnd() % 152);
        std::vector<uint8_t> b(l + 2);
        /* sequence: random, with planted homopolymers / short-period repeats / noisy repeats */
        for (int i = 0; i < l; ) {
            const int kind = rnd() % 10;
            if (kind < 5) { b[i++] = rnd() % 4; continue; }
            const int p = 1 + rnd() % 8, copies = 2 + rnd() % 12;
            uint8_t unit[8]; for (int k = 0; k < p; k++) unit[k] = rnd() % 4;
            if (kind == 9) for (int k = 1; k < p; k++) unit[k] = unit[0];       /* unit that is itself a repeat */
            for (int c = 0; c < copies * p && i < l; c++) { b[i] = unit[c % p]; if (kind >= 7 && rnd() % 23 == 0) b[i] = rnd() % 4; i++; }
        }
        if (rnd() % 16 == 0) for (int i = 0; i < l; i++) if (rnd() % 50 == 0) b[i] = 4;   /* N and other codes map to 'A' (str_finder.c:15-32) */
        memset(seq4.data(), 0, (size_t)(l + 3) / 2 + 1);
        for (int i = 0; i < l; i++) { const uint8_t code = b[i] == 4 ? 15 : nt[b[i]]; seq4[i >> 1] |= (i & 1) ? code : code << 4; }
        const int phantom = rnd() % 16;
        /* CIGAR: soft clips, M, a few indels; query length l */
        uint32_t cig[16]; int nc = 0, left = l;
        if (rnd() % 4 == 0 && left > 20) { int s = 1 + rnd() % 10; cig[nc++] = s << 4 | 4; left -= s; }
        while (left > 0 && nc < 12) {
            int m = (rnd() % 3 == 0 && left > 30) ? 1 + rnd() % (left - 1) : left;
            cig[nc++] = m << 4 | 0; left -= m;
            if (left > 0) { const int r = rnd() % 3; if (r == 0) cig[nc++] = (1 + rnd() % 5) << 4 | 2; else if (r == 1) { int x = 1 + rnd() % 4; if (x > left) x = left; cig[nc++] = x << 4 | 1; left -= x; } else cig[nc++] = (1 + rnd() % 300) << 4 | 3; }
        }
        if (left > 0) cig[nc++] = left << 4 | 4;
        const int read_pos = 1000 + rnd() % 100000;
        for (int rep = 0; rep < 3; rep++) {
            const int rpos = 1 + rnd() % l, add = rnd() % 7, pos = read_pos + rnd() % 200;
            int lo1 = pos, hi1 = pos, lo2 = pos, hi2 = pos;
            cg_mask_lc(seq4.data(), l, phantom, cig, nc, read_pos, rpos, add, win.data(), L, &lo1, &hi1);
            if (L->overflow) continue;
            int16_t live[16];
            cg_mask_lc_lean<1>(seq4.data(), l, phantom, cig, nc, read_pos, rpos, add, live, &lo2, &hi2);
            int lo3 = pos, hi3 = pos;
            uint64_t W[16]; uint32_t ring[16];
            cg_mask_lc_bits<1, 1>(seq4.data(), l, phantom, cig, nc, read_pos, rpos, add, W, ring, &lo3, &hi3);
            if (lo1 != pos || hi1 != pos) nonempty++;
            if (lo1 != lo3 || hi1 != hi3) {
                if (bad++ < 10) {
                    fprintf(stderr, "MISMATCH(bits) l=%d rpos=%d add=%d list [%d,%d] bits [%d,%d]\n  ", l, rpos, add, lo1 - pos, hi1 - pos, lo3 - pos, hi3 - pos);
                    for (int i = 0; i < l && i < 320; i++) fputc("ACGTN"[b[i]], stderr);
                    fputc('\n', stderr);
                }
            }
            if (lo1 != lo2 || hi1 != hi2) {
                if (bad++ < 10) {
                    fprintf(stderr, "MISMATCH l=%d rpos=%d add=%d list [%d,%d] lean [%d,%d]\n  ", l, rpos, add, lo1, hi1, lo2, hi2);
                    for (int i = 0; i < l && i < 320; i++) fputc("ACGTN"[b[i]], stderr);
                    fputc('\n', stderr);
                }
            }
        }
    }
    printf("cases=%ld nonempty=%ld mismatches=%ld\n", N * 3, nonempty, bad);
    return bad != 0;
}
