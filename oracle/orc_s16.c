/*
 * oracle/orc_s16.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A CPU model of the NEXT fill kernel's arithmetic: the linear GuidedAlign recurrence
 * (GuidedAlign.h:474-624, tie order Diagonal > Left > Up :560-568) swept by anti-diagonals over per-diagonal slots, the
 * way bgpu_fill.cu does it, but with every slot held as a 16-bit value (score << 2 | arrow) RELATIVE to a per-job
 * offset that is re-based every 64 anti-diagonals -- the representation two jobs per lane in VIADDMNMX.S16x2 halves
 * would use.  The same sweep is run in 32 bits without re-basing (today's kernel arithmetic); the model reports whether
 * every in-band cell gets the same arrow and the same score in both, and how much of the 16-bit range was used.
 * It proves (or refutes) the number format before any device code is written; the 32-bit side is itself pinned by
 * comparing its end score with orc_align() in tests/test_s16_model.py.
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "orc_align.h"

#define BIG32 (1 << 30)
#define BIG16 24000            /* out-of-band / unreachable, shifted domain; legit relative values stay below THR16 */
#define THR16 20000
enum { TG_DIAG = 0, TG_LEFT = 1, TG_UP = 2, TG_NONE = 3 };

static int code_of(uint8_t c) {                             /* ThreeBit, NucConversion.h:48-84 (0..4, else -1) */
  if (c <= 4) return c;
  switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
    case 'N': case 'n': return 4;
    default: return -1;
  }
}

/* out[0] cells compared, [1] arrow mismatches, [2] score mismatches, [3] end score (32-bit sweep), [4] end score
 * (16-bit sweep + offset), [5] min and [6] max legit relative value seen (shifted domain), [7] max BIG-ish value seen,
 * [8] number of re-bases.  Returns 0, -1 on unsupported input (non-ACGTN bases, empty guide). */
int orc_guided_s16_model(const orc_scorefn *fn, const orc_job *job, int64_t *out) {
  const uint32_t capRows = job->qLen + 2;
  int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)capRows);
  int64_t nCellsRef = 0;
  const int nRows = orc_guide_rows(job->guide, job->nGuide, job->band, rows, capRows, &nCellsRef);
  memset(out, 0, sizeof(int64_t) * 9);
  if (nRows <= 1) { free(rows); return -1; }
  const int Qn = nRows - 1;                                   /* guide rows 1..Qn, row 0 = boundary row */
  const int qStart = rows[4 + 0], tStart = rows[4 + 1];
  const int tEnd = rows[4 * Qn + 1] + 1, Tn = tEnd - tStart;
  /* per row: in-band columns t' in [lo, hi] (t' = t - tStart + 1), clipped as the fill does (t < tEnd, :502-503) */
  int *lo = (int *)malloc(sizeof(int) * (size_t)nRows), *hi = (int *)malloc(sizeof(int) * (size_t)nRows);
  int i;
  for (i = 0; i < nRows; i++) {
    const int t = rows[4 * i + 1], pre = rows[4 * i + 2], post = rows[4 * i + 3];
    lo[i] = t - pre - tStart + 1; hi[i] = t + post - tStart + 1;
    if (hi[i] > Tn) hi[i] = Tn;
    if (i == 0) { lo[i] = 0; }
  }
  const int global = job->alignType == ORC_GLOBAL;
  const int del = fn->del, ins = fn->ins;
  /* slots by diagonal cd = t' - q' + Qn in [0, Qn + Tn] */
  const int nDiag = Qn + Tn + 1;
  int *s32 = (int *)malloc(sizeof(int) * (size_t)nDiag);
  int16_t *s16 = (int16_t *)malloc(sizeof(int16_t) * (size_t)nDiag);
  for (i = 0; i < nDiag; i++) { s32[i] = BIG32 | TG_NONE; s16[i] = (int16_t)(BIG16 | TG_NONE); }
  int64_t offset = 0;                                         /* shifted domain: value32 = value16 + offset */
  int64_t cells = 0, badArrow = 0, badScore = 0, minRel = INT_MAX, maxRel = INT_MIN, maxBig = 0, rebases = 0;
  int end32 = 0; int64_t end16 = 0;
  const int delT = (del << 2) | TG_LEFT, insT = (ins << 2) | TG_UP;
  int d;
  for (d = 0; d <= Qn + Tn; d++) {
    if (d > 0 && (d & 63) == 0) {
      /* re-base: the smallest legit slot becomes 0; BIG-ish slots are reset to BIG16 */
      int base = INT_MAX;
      for (i = 0; i < nDiag; i++) if (s16[i] < THR16 && (s16[i] & ~3) < base) base = s16[i] & ~3;
      if (base != INT_MAX) {
        for (i = 0; i < nDiag; i++) {
          if (s16[i] < THR16) s16[i] = (int16_t)(s16[i] - base);
          else s16[i] = (int16_t)(BIG16 | (s16[i] & 3));
        }
        offset += base; rebases++;
      }
    }
    /* cells of anti-diagonal d: q' from max(0, d - Tn) to min(Qn, d) */
    int qlo = d - Tn; if (qlo < 0) qlo = 0;
    int qhi = d < Qn ? d : Qn;
    /* two passes keep the neighbours of this anti-diagonal intact: compute into temporaries first */
    int nq = qhi - qlo + 1;
    if (nq <= 0) continue;
    int *n32 = (int *)malloc(sizeof(int) * (size_t)nq);
    int16_t *n16 = (int16_t *)malloc(sizeof(int16_t) * (size_t)nq);
    int q;
    for (q = qlo; q <= qhi; q++) {
      const int t = d - q, cd = t - q + Qn, k = q - qlo;
      const int inb = t >= lo[q] && t <= hi[q];
      int v32, v16;
      if (!inb) { v32 = BIG32 | TG_NONE; v16 = BIG16 | TG_NONE; }
      else if (q == 0) {                                      /* boundary row :415-442 */
        v32 = ((t * (global ? del : 0)) << 2) | TG_LEFT;
        v16 = (int)((int64_t)v32 - offset);
      } else {
        const int qc = code_of(job->q[qStart + q - 1]);
        const int tc = t >= 1 ? code_of(job->t[tStart + t - 1]) : 0;
        if (qc < 0 || tc < 0) { free(n32); free(n16); free(rows); free(lo); free(hi); free(s32); free(s16); return -1; }
        const int m = t >= 1 ? fn->M[qc * 5 + tc] << 2 : 0;
        /* neighbours: same diagonal (q-1,t-1), diagonal-1 (q,t-1), diagonal+1 (q-1,t); a neighbour outside the matrix is BIG */
        const int dg32 = t >= 1 ? s32[cd] & ~3 : BIG32, lf32 = (t >= 1 && cd >= 1) ? s32[cd - 1] & ~3 : BIG32, up32 = cd + 1 < nDiag ? s32[cd + 1] & ~3 : BIG32;
        const int dg16 = t >= 1 ? s16[cd] & ~3 : BIG16, lf16 = (t >= 1 && cd >= 1) ? s16[cd - 1] & ~3 : BIG16, up16 = cd + 1 < nDiag ? s16[cd + 1] & ~3 : BIG16;
        int c32 = dg32 + m, c16 = dg16 + m;                   /* Diagonal, tag 0 */
        if (lf32 + delT < c32) c32 = lf32 + delT;
        if (up32 + insT < c32) c32 = up32 + insT;
        if (lf16 + delT < c16) c16 = lf16 + delT;
        if (up16 + insT < c16) c16 = up16 + insT;
        v32 = c32; v16 = c16;
        if ((v32 & ~3) >= BIG32 / 2) v32 = (v32 & ~3) | TG_NONE;   /* unreachable: the reference leaves NoArrow */
        if (v16 >= THR16) v16 = (v16 & ~3) | TG_NONE;
      }
      n32[k] = v32; n16[k] = (int16_t)v16;
      if (v16 > 32767 || v16 < -32768) badScore++;            /* would not fit: counted, the run continues */
      if (inb) {
        cells++;
        const int reach = (v32 & ~3) < BIG32 / 2;
        if (reach) {
          if ((v32 & 3) != (v16 & 3)) badArrow++;
          if ((int64_t)(v32 & ~3) != (int64_t)(v16 & ~3) + offset) badScore++;
          if ((v16 & ~3) < minRel) minRel = v16 & ~3;
          if ((v16 & ~3) > maxRel) maxRel = v16 & ~3;
          if (v16 >= THR16) badScore++;                       /* a legit value crossed into the BIG zone */
        } else {
          if (v16 < THR16) badScore++;                        /* an unreachable cell looks legit in 16 bits */
          if (v16 > maxBig) maxBig = v16;
        }
        if (q == Qn && t == Tn) { end32 = v32 >> 2; end16 = ((int64_t)(v16 & ~3) + offset) >> 2; }
      }
    }
    for (q = qlo; q <= qhi; q++) { const int cd = (d - q) - q + Qn; s32[cd] = n32[q - qlo]; s16[cd] = n16[q - qlo]; }
    free(n32); free(n16);
  }
  out[0] = cells; out[1] = badArrow; out[2] = badScore; out[3] = end32; out[4] = end16;
  out[5] = minRel == INT_MAX ? 0 : minRel; out[6] = maxRel == INT_MIN ? 0 : maxRel; out[7] = maxBig; out[8] = rebases;
  free(rows); free(lo); free(hi); free(s32); free(s16);
  return 0;
}
