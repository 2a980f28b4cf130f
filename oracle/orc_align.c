/*
 * oracle/orc_align.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's per-candidate refinement path.  It is the
 * checker the CUDA path is compared against; it must never be the thing measured or
 * shipped.  Each function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Parity status: the reference holds no golden vectors
 * for this path, so this file is pinned against the reference code
 * itself compiled into oracle/_ref/libblasr_ref.so (tests/test_oracle_vs_ref.py) and
 * against fixtures generated from that library (tests/golden/, made by
 * tests/golden/make_golden.py).
 *
 * Restated semantics, not copied text: cells are addressed by (row, absolute column)
 * with explicit per-row ranges instead of the reference's sticky buffer index; quirks
 * that change results are kept and called out where they occur.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <pthread.h>
#include "orc_align.h"

#define ORC_INF INT_MAX /* defs.h:15-17 */
#define MAX_BAND 250    /* GuidedAlign.h:29 */

/* path codes (datastructures/alignment/Path.h:4-19), only the ones this path produces */
enum { A_DIAG = 0, A_UP = 1, A_LEFT = 2, A_INS_UP = 3, A_INS_OPEN = 4, A_INS_CLOSE = 5,
       A_DEL_LEFT = 6, A_DEL_OPEN = 7, A_DEL_CLOSE = 8, A_NONE = 12 };

/* ---- base codes: NucConversion.h:48-84 (ThreeBit) ------------------------------ */
static uint8_t g_code[256];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_codes(void) {
  memset(g_code, 255, sizeof g_code);
  const char *acgt = "ACGT";
  for (int i = 0; i < 4; i++) {
    g_code[(unsigned char)acgt[i]] = (uint8_t)i;
    g_code[(unsigned char)(acgt[i] + 32)] = (uint8_t)i;
    g_code[i] = (uint8_t)i;
  }
  g_code[4] = 4;
  g_code['$'] = 5;
  /* IUPAC ambiguity letters collapse onto N (=4); note the table's asymmetries:
   * 'X' is unmapped while 'x' and 'y' are N (NucConversion.h:62-69). */
  const char *upper = "BDHKMNRSUVWY";
  for (const char *p = upper; *p; p++) { g_code[(unsigned char)*p] = 4; g_code[(unsigned char)(*p + 32)] = 4; }
  g_code['x'] = 4;
}
static inline int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); } /* wrap like -O3 x86 */
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int iabs(int a) { return a < 0 ? -a : a; }

/* ---- score functions ---------------------------------------------------------- */
/* DistanceMatrixScoreFunction<DNASequence,FASTQSequence>::Match, DistanceMatrixScoreFunction.h:100-105
 * (row = query code); QualityValueScoreFunction::Match, QualityValueScoreFunction.h:78-83 with
 * QVDistanceMatrix (ScoreMatrices.h:4-10): -qual on equal ACGT, +qual otherwise (N-N is +qual). */
static inline int match_cost(const orc_scorefn *fn, const orc_job *j, uint32_t tpos, uint32_t qpos) {
  if (fn->kind == ORC_FN_IDS) {
    /* IDSScoreFunction<DNASequence,FASTQSequence>::Match, IDSScoreFunction.h:126-139: raw (case-sensitive) bytes */
    if (j->q[qpos] == j->t[tpos]) return 0;
    if (j->subTag[qpos] == j->t[tpos]) return (int)j->subQV[qpos];
    return fn->substitutionPrior;
  }
  int qc = g_code[j->q[qpos]], tc = g_code[j->t[tpos]];
  if (fn->kind == ORC_FN_QUALITY) {
    int sign = (qc == tc && qc < 4) ? -1 : 1;
    return sign * (int)j->qual[qpos];
  }
  return fn->M[qc * 5 + tc];
}
/* the 4-argument Insertion / Deletion the DP loops call.  Distance / QualityValue: the constants ins / del
 * (DistanceMatrixScoreFunction.h:34-45, QualityValueScoreFunction.h:47-66).  IDS: insertionQV[q]
 * (IDSScoreFunction.h:113-116) and the deletion-tag rule (:80-103; the mergeQV branch is `if (false)`). */
static inline int ins_cost(const orc_scorefn *fn, const orc_job *j, uint32_t qpos) {
  if (fn->kind == ORC_FN_IDS) return (int)j->insQV[qpos];
  return fn->ins;
}
static inline int del_cost(const orc_scorefn *fn, const orc_job *j, uint32_t tpos, uint32_t qpos) {
  if (fn->kind == ORC_FN_IDS) {
    if (j->delQV && j->delTag) {
      if (j->delTag[qpos] == 'N') return fn->globalDeletionPrior;
      if (j->delTag[qpos] == j->t[tpos]) return (int)j->delQV[qpos];
      return fn->globalDeletionPrior;
    }
    return fn->del;
  }
  return fn->del;
}

/* ---- growable output path ------------------------------------------------------ */
typedef struct { uint8_t *a; size_t n, cap; } path_t;
static void path_push(path_t *p, uint8_t v) {
  if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 1024; p->a = (uint8_t *)realloc(p->a, p->cap); }
  p->a[p->n++] = v;
}
static void path_reverse(path_t *p) {
  for (size_t i = 0, k = p->n; i + 1 < k; i++, k--) { uint8_t x = p->a[i]; p->a[i] = p->a[k - 1]; p->a[k - 1] = x; }
}

/* ---- alignment container ------------------------------------------------------- */
typedef struct {
  uint32_t qPos, tPos;
  uint32_t *blocks; uint32_t nBlocks, capBlocks;          /* {q,t,len} */
  uint32_t *gapCounts; uint32_t nGapLists, capGapLists;
  int32_t *gaps; uint32_t nGaps, capGaps;                /* {seq,len} */
  int overflow;
} aln_t;

static void aln_push_block(aln_t *a, uint32_t q, uint32_t t, uint32_t len) {
  if (a->nBlocks < a->capBlocks) {
    a->blocks[3 * a->nBlocks] = q; a->blocks[3 * a->nBlocks + 1] = t; a->blocks[3 * a->nBlocks + 2] = len;
  } else a->overflow = 1;
  a->nBlocks++;
}
static void aln_push_gaplist(aln_t *a) {
  if (a->nGapLists < a->capGapLists) a->gapCounts[a->nGapLists] = 0; else a->overflow = 1;
  a->nGapLists++;
}
static void aln_push_gap(aln_t *a, int seq, int len) {
  if (a->nGaps < a->capGaps) { a->gaps[2 * a->nGaps] = seq; a->gaps[2 * a->nGaps + 1] = len; } else a->overflow = 1;
  a->nGaps++;
  if (a->nGapLists <= a->capGapLists && a->nGapLists > 0) a->gapCounts[a->nGapLists - 1]++;
}

/* Alignment::ArrowPathToAlignment, datastructures/alignment/Alignment.h:190-254.
 * Forward path -> blocks + one gap list before the first block and after every block;
 * the list after the last block is emptied.  Left -> Gap::Query(0), Up -> Gap::Target(1). */
static void arrows_to_alignment(aln_t *a, const uint8_t *p, size_t n) {
  size_t i = 0; uint32_t q = 0, t = 0; int first = 1;
  while (i < n) {
    if (!first && p[i] == A_DIAG) {
      uint32_t len = 0, bq = q, bt = t;
      while (i < n && p[i] == A_DIAG) { len++; i++; q++; t++; }
      aln_push_block(a, bq, bt, len);
    }
    aln_push_gaplist(a);
    uint32_t listStartGap = a->nGaps;
    while (i < n && (p[i] == A_LEFT || p[i] == A_UP)) {
      uint8_t kind = p[i]; size_t s = i;
      while (i < n && p[i] == kind) { i++; if (kind == A_LEFT) t++; else q++; }
      aln_push_gap(a, kind == A_LEFT ? 0 : 1, (int)(i - s));
    }
    if (i == n) { /* trailing list is cleared (Alignment.h:246-250) */
      a->nGaps = listStartGap;
      if (a->nGapLists <= a->capGapLists) a->gapCounts[a->nGapLists - 1] = 0;
    }
    if (i < n && !first && p[i] != A_DIAG) break; /* unreachable for paths this file builds */
    first = 0;
  }
}

/* RemoveAlignmentPrefixGaps, AlignmentUtils.h:620-644 */
static void remove_prefix_gaps(aln_t *a) {
  if (a->nGapLists == 0) return;
  uint32_t n0 = a->gapCounts[0], qs = 0, ts = 0;
  for (uint32_t g = 0; g < n0 && g < a->capGaps; g++) {
    if (a->gaps[2 * g] == 1) qs += (uint32_t)a->gaps[2 * g + 1]; else ts += (uint32_t)a->gaps[2 * g + 1];
  }
  uint32_t nb = a->nBlocks < a->capBlocks ? a->nBlocks : a->capBlocks;
  for (uint32_t b = 0; b < nb; b++) { a->blocks[3 * b] -= qs; a->blocks[3 * b + 1] -= ts; }
  if (n0) {
    uint32_t ng = a->nGaps < a->capGaps ? a->nGaps : a->capGaps;
    memmove(a->gaps, a->gaps + 2 * n0, (size_t)(ng - n0) * 2 * sizeof(int32_t));
    a->nGaps -= n0; a->gapCounts[0] = 0;
  }
  a->tPos += ts; a->qPos += qs;
}

/* ---- guide construction: AlignmentToGuide, GuidedAlign.h:104-259 ----------------- */
typedef struct { int q, t, tPre, tPost; } grow_t;

/* ComputeDrift(Block,Block), AlignmentUtils.h:586-600: the common-gap shuffle cancels,
 * the value is always tGap - qGap. */
static int block_drift(const uint32_t *cur, const uint32_t *next) {
  int tGap = (int)(next[1] - (cur[1] + cur[2])), qGap = (int)(next[0] - (cur[0] + cur[2]));
  return tGap - qGap;
}

static int build_guide(const uint32_t *g, uint32_t n, int band, grow_t **out) {
  *out = NULL;
  if (n == 0) return 0;
  int tStart = (int)g[1], qStart = (int)g[0];
  int qEnd = (int)(g[3 * (n - 1)] + g[3 * (n - 1) + 2]);
  int nRows = qEnd - qStart + 1;
  if (nRows < 1) return -1;
  grow_t *r = (grow_t *)calloc((size_t)nRows, sizeof(grow_t));
  r[0].t = tStart - 1; r[0].q = qStart - 1;                       /* :126-127 */
  int drift = iabs(tStart - qStart);                              /* :128 */
  r[0].tPost = drift > band ? drift : band; r[0].tPre = 0;         /* :129-135 */
  int gi = 1;
  for (uint32_t b = 0; b < n; b++) {
    const uint32_t *blk = g + 3 * b;
    for (uint32_t bp = 0; bp < blk[2]; bp++, gi++) {
      if (gi >= nRows) { free(r); return -1; }
      r[gi].t = (int)(blk[1] + bp); r[gi].q = (int)(blk[0] + bp);
      int reach = r[gi].t - (r[gi - 1].t - r[gi - 1].tPre);       /* back to previous left edge */
      if (bp == 0) { r[gi].tPre = reach; r[gi].tPost = band + iabs(drift); }   /* :161-165, unclamped */
      else { r[gi].tPre = imin(band, reach); r[gi].tPost = imin(MAX_BAND, band); } /* :170-174 */
    }
    if (b + 1 < n) {
      const uint32_t *nx = blk + 3;
      int qGap = (int)(nx[0] - (blk[0] + blk[2])), tGap = (int)(nx[1] - (blk[1] + blk[2]));
      drift = block_drift(blk, nx);                                /* :199 */
      int diag = imin(qGap, tGap);
      int qp = (int)(blk[0] + blk[2]), tp = (int)(blk[1] + blk[2]), qe = (int)nx[0];
      for (int d = 0; d < diag; d++, tp++, qp++, gi++) {            /* :213-223 */
        if (gi >= nRows) { free(r); return -1; }
        r[gi].t = tp; r[gi].q = qp;
        r[gi].tPre = imin(MAX_BAND, r[gi].t - (r[gi - 1].t - r[gi - 1].tPre));
        r[gi].tPost = imin(MAX_BAND, band + iabs(drift));
      }
      while (qp < qe) {                                            /* :239-250, t frozen */
        if (gi >= nRows) { free(r); return -1; }
        r[gi].t = tp; r[gi].q = qp; qp++;
        r[gi].tPre = imin(MAX_BAND, r[gi].t - (r[gi - 1].t - r[gi - 1].tPre));
        r[gi].tPost = imin(MAX_BAND, band + iabs(drift));
        gi++;
      }
    }
  }
  if (gi != nRows) { free(r); return -1; }
  *out = r;
  return nRows;
}

int orc_guide_rows(const uint32_t *guide, uint32_t nGuide, int band, int32_t *rows, uint32_t capRows,
                   int64_t *nCells) {
  grow_t *r; int n = build_guide(guide, nGuide, band, &r);
  if (n <= 0) { if (nCells) *nCells = 0; return n < 0 ? -2 : 0; }
  int64_t c = 0;
  for (int i = 0; i < n; i++) c += r[i].tPre + r[i].tPost + 1;     /* ComputeMatrixNElem :83-92 */
  if (nCells) *nCells = c;
  if ((uint32_t)n > capRows) { free(r); return -1; }
  for (int i = 0; i < n; i++) { rows[4 * i] = r[i].q; rows[4 * i + 1] = r[i].t; rows[4 * i + 2] = r[i].tPre; rows[4 * i + 3] = r[i].tPost; }
  free(r);
  return n;
}

/* ---- GuidedAlign (GuidedAlign.h:278-685) and AffineGuidedAlign (AffineGuidedAlign.h:31-488) ----
 * Row i (i = q - qStart + 1) holds absolute columns [lo_i, hi_i] = [t-tPre, t+tPost], stored
 * contiguously (StoreMatrixOffsets :94-101).  A neighbour contributes only if its column lies in
 * its row's range (GetBufferIndexFunctor :45-79; the sticky fast path is equivalent to the range
 * test because left edges never move left). */
typedef struct {
  int n; grow_t *r; int64_t *off; /* off[i] = index of column lo_i */
  int *S, *AI, *AD; uint8_t *P, *PI, *PD;
} gmat_t;

static inline int64_t gidx(const gmat_t *m, int row, int col) {
  if (row < 0 || row >= m->n) return -1;
  const grow_t *g = &m->r[row];
  if (col < g->t - g->tPre || col > g->t + g->tPost) return -1;
  return m->off[row] + (col - (g->t - g->tPre));
}

static int guided_align(const orc_scorefn *fn, const orc_job *j, int affine, orc_result *res, aln_t *aln) {
  gmat_t m; memset(&m, 0, sizeof m);
  m.n = build_guide(j->guide, j->nGuide, j->band, &m.r);
  if (m.n == 0) { res->status = ORC_EMPTY_GUIDE; res->score = 0; return 0; } /* :388-392 */
  if (m.n < 0) { res->status = ORC_BAD_INPUT; return 0; }
  int64_t nCells = 0;
  m.off = (int64_t *)malloc(sizeof(int64_t) * (size_t)m.n);
  for (int i = 0; i < m.n; i++) {
    if (m.r[i].tPre < 0 || m.r[i].tPost < 0) { res->status = ORC_BAD_INPUT; free(m.r); free(m.off); return 0; }
    m.off[i] = nCells; nCells += m.r[i].tPre + m.r[i].tPost + 1;
  }
  if (nCells > INT_MAX) { res->status = ORC_BAD_INPUT; free(m.r); free(m.off); return 0; }
  int qStart = m.r[1].q, tStart = m.r[1].t;
  int qEnd = m.r[m.n - 1].q + 1, tEnd = m.r[m.n - 1].t + 1;
  if ((uint32_t)qEnd > j->qLen || (uint32_t)tEnd > j->tLen) { res->status = ORC_BAD_INPUT; free(m.r); free(m.off); return 0; }
  m.S = (int *)calloc((size_t)nCells, sizeof(int));                 /* zero / NoArrow fill :375-376 */
  m.P = (uint8_t *)malloc((size_t)nCells); memset(m.P, A_NONE, (size_t)nCells);
  if (affine) {                                                     /* AffineGuidedAlign.h:107-116 */
    m.AI = (int *)malloc(sizeof(int) * (size_t)nCells); m.AD = (int *)malloc(sizeof(int) * (size_t)nCells);
    for (int64_t i = 0; i < nCells; i++) m.AI[i] = m.AD[i] = fn->affineOpen;
    m.PI = (uint8_t *)malloc((size_t)nCells); memset(m.PI, A_NONE, (size_t)nCells);
    m.PD = (uint8_t *)malloc((size_t)nCells); memset(m.PD, A_NONE, (size_t)nCells);
  }
  int global = (j->alignType == ORC_GLOBAL), local = (j->alignType == ORC_LOCAL);
  /* boundary row :415-442 (every column of row 0 right of the origin, even past tEnd) */
  for (int t = tStart; t < tStart + m.r[0].tPost; t++) {
    int64_t c = gidx(&m, 0, t), d = gidx(&m, 0, t - 1);
    if (c < 0) { res->status = ORC_BAD_INPUT; goto done; }
    if (d >= 0) {
      if (global) m.S[c] = wadd(m.S[d], fn->del); else if (local) m.S[c] = 0;
      m.P[c] = A_LEFT;
      if (affine) { m.PD[c] = A_DEL_OPEN; m.PI[c] = A_INS_OPEN; }
    }
  }
  /* left stripe :447-470 -- every one of these cells is recomputed by the fill below
   * (column tStart-1 >= -1 is never skipped), so it only matters for rows the fill never visits. */
  for (int q = qStart; q < qStart + j->band && q < qEnd; q++) {
    int64_t u = gidx(&m, q - qStart, tStart - 1), c = gidx(&m, q - qStart + 1, tStart - 1);
    if (u >= 0 && c >= 0) {
      m.S[c] = global ? wadd(m.S[u], fn->ins) : 0;
      m.P[c] = A_UP;
      if (affine) { m.PI[c] = A_INS_OPEN; m.PD[c] = A_DEL_OPEN; }
    }
  }
  /* fill :474-624 / AffineGuidedAlign.h:241-375 */
  for (int q = qStart; q < qEnd; q++) {
    int row = q - qStart + 1;
    int lo = m.r[row].t - m.r[row].tPre, hi = m.r[row].t + m.r[row].tPost;
    for (int t = lo; t <= hi; t++) {
      if (t < -1 || t >= tEnd) continue;                             /* :502-503 */
      int64_t c = gidx(&m, row, t), dg = gidx(&m, row - 1, t - 1), up = gidx(&m, row - 1, t), lf = gidx(&m, row, t - 1);
      int ms = dg >= 0 ? wadd(m.S[dg], match_cost(fn, j, (uint32_t)t, (uint32_t)q)) : ORC_INF;
      int is = up >= 0 ? wadd(m.S[up], ins_cost(fn, j, (uint32_t)q)) : ORC_INF;
      int ds = lf >= 0 ? wadd(m.S[lf], del_cost(fn, j, (uint32_t)t, (uint32_t)q)) : ORC_INF;
      if (!affine) {
        int best = imin(ms, imin(is, ds));
        m.S[c] = best;
        /* tie order Diagonal > Left > Up :560-568 */
        m.P[c] = best == ORC_INF ? A_NONE : best == ms ? A_DIAG : best == ds ? A_LEFT : A_UP;
      } else {
        int ie = up >= 0 ? wadd(m.AI[up], fn->affineExtend) : ORC_INF;
        int de = lf >= 0 ? wadd(m.AD[lf], fn->affineExtend) : ORC_INF;
        int best = imin(ms, imin(is, imin(ds, imin(ie, de))));
        m.S[c] = best;
        /* tie order Diagonal > Left > Up > AffineInsClose > AffineDelClose, AffineGuidedAlign.h:326-341 */
        m.P[c] = best == ORC_INF ? A_NONE : best == ms ? A_DIAG : best == ds ? A_LEFT : best == is ? A_UP
                 : best == ie ? A_INS_CLOSE : A_DEL_CLOSE;
        int open = wadd(best, fn->affineOpen);
        if (open < ie) { m.PI[c] = A_INS_OPEN; m.AI[c] = open; } else { m.PI[c] = A_INS_UP; m.AI[c] = ie; }   /* :357-364 */
        if (open < de) { m.PD[c] = A_DEL_OPEN; m.AD[c] = open; } else { m.PD[c] = A_DEL_LEFT; m.AD[c] = de; } /* :366-373 */
      }
    }
  }
  /* traceback :626-663 / AffineGuidedAlign.h:377-468 */
  {
    path_t p = {0, 0, 0};
    int q = qEnd - 1, t = tEnd - 1, mat = 0; /* 0 match, 1 affine ins, 2 affine del */
    while (q >= qStart || t >= tStart) {
      int64_t c = gidx(&m, q - qStart + 1, t);
      if (c < 0) { res->status = ORC_PATH_AWRY; break; }
      if (mat == 0) {
        uint8_t a = m.P[c];
        if (a == A_NONE) { res->status = ORC_PATH_AWRY; break; }   /* reference: exit(1) */
        if (a == A_DIAG) { path_push(&p, A_DIAG); q--; t--; }
        else if (a == A_UP) { path_push(&p, A_UP); q--; }
        else if (a == A_LEFT) { path_push(&p, A_LEFT); t--; }
        else if (a == A_INS_CLOSE) { path_push(&p, A_UP); mat = 1; q--; }
        else if (a == A_DEL_CLOSE) { path_push(&p, A_LEFT); mat = 2; t--; }
      } else if (mat == 1) {
        uint8_t a = m.PI[c];
        if (a == A_INS_OPEN) mat = 0;
        else if (a == A_INS_UP) { q--; path_push(&p, A_UP); }
        else { res->status = ORC_PATH_AWRY; break; }                /* reference: assert(0) */
      } else {
        uint8_t a = m.PD[c];
        if (a == A_DEL_OPEN) mat = 0;
        else if (a == A_DEL_LEFT) { t--; path_push(&p, A_LEFT); }
        else { res->status = ORC_PATH_AWRY; break; }
      }
    }
    if (res->status == ORC_OK) {
      path_reverse(&p);
      aln->qPos = (uint32_t)qStart; aln->tPos = (uint32_t)tStart;   /* :667-668 */
      arrows_to_alignment(aln, p.a, p.n);
      remove_prefix_gaps(aln);
      res->nCells = (int)nCells;
      int64_t last = gidx(&m, qEnd - qStart, tEnd - 1);
      res->score = res->alnScore = last >= 0 ? m.S[last] : 0;        /* :675-684 */
    }
    free(p.a);
  }
done:
  free(m.r); free(m.off); free(m.S); free(m.P); free(m.AI); free(m.AD); free(m.PI); free(m.PD);
  return 0;
}

/* ---- KBandAlign, KBandAlign.h:75-403 ------------------------------------------- */
/* SetKBoundedLengths :36-56 */
static void kbounded(uint32_t tLength, uint32_t qLength, uint32_t k, uint32_t *tLen, uint32_t *qLen) {
  if (tLength < qLength) { *tLen = tLength; *qLen = qLength < tLength + k ? qLength : tLength + k; }
  else if (qLength < tLength) { *qLen = qLength; *tLen = tLength < qLength + k ? tLength : qLength + k; }
  else { *qLen = qLength; *tLen = tLength; }
}

static int kband_align(const orc_scorefn *fn, const orc_job *j, orc_result *res, aln_t *aln) {
  int k = j->band, at = j->alignType;
  if (k < 0) { res->status = ORC_BAD_INPUT; return 0; }
  uint32_t tLen, qLen; kbounded(j->tLen, j->qLen, (uint32_t)k, &tLen, &qLen);
  /* TargetFit/Fit with k > tLen reads an uninitialised index in the reference (:286,:307) */
  if ((at == ORC_TARGETFIT || at == ORC_FIT) && (uint32_t)k > tLen) { res->status = ORC_BAD_INPUT; return 0; }
  int64_t nCols = 2 * (int64_t)k + 1, total = ((int64_t)qLen + 1) * nCols;
  if (total > INT_MAX) { res->status = ORC_BAD_INPUT; return 0; }
  res->nCells = (int)total;                                          /* :97-98 */
  int *S = (int *)calloc((size_t)total, sizeof(int));
  uint8_t *P = (uint8_t *)malloc((size_t)total); memset(P, A_NONE, (size_t)total);
#define KB(q_, c_) ((int64_t)(q_) * nCols + (c_))                    /* band column c = k + t - q */
  for (int q = 1; q <= k && q < (int64_t)qLen + 1; q++) { S[KB(q, k - q)] = q * j->bndIns; P[KB(q, k - q)] = A_UP; } /* :115-118 */
  if (at == ORC_GLOBAL)
    for (int t = 1; t <= k && (uint32_t)t < tLen; t++) { S[KB(0, t + k)] = t * j->bndDel; P[KB(0, t + k)] = A_LEFT; } /* :119-124 */
  if (at == ORC_QUERYFIT || at == ORC_FIT)
    for (int t = 1; t <= k && (uint32_t)t < tLen; t++) { S[KB(0, t + k)] = 0; P[KB(0, t + k)] = A_LEFT; }           /* :125-130 */
  if (at == ORC_TARGETFIT || at == ORC_FIT)
    for (int q = 1; q <= k && (uint32_t)q < qLen; q++) { S[KB(q, 0)] = 0; P[KB(q, 0)] = A_UP; }  /* :131-136: band column 0, kept as is */
  S[KB(0, k)] = 0; P[KB(0, k)] = A_DIAG;                             /* :141-142 */
  for (int q = 1; q <= (int)qLen; q++) {
    for (int t = q - k; t < q + k + 1; t++) {
      if (t < 1 || (uint32_t)t > tLen) continue;
      int ds = (t == q - k) ? ORC_INF : wadd(S[KB(q, k + t - q - 1)], del_cost(fn, j, (uint32_t)t - 1, (uint32_t)q - 1));        /* :155-164 */
      int ms = wadd(S[KB(q - 1, k + t - q)], match_cost(fn, j, (uint32_t)t - 1, (uint32_t)q - 1)); /* :176-177 */
      int is = (t == q + k) ? ORC_INF : wadd(S[KB(q - 1, k + t - q + 1)], ins_cost(fn, j, (uint32_t)q - 1));   /* :182-190 */
      int best = imin(ms, imin(is, ds));
      S[KB(q, k + t - q)] = best;
      P[KB(q, k + t - q)] = best == ms ? A_DIAG : best == ds ? A_LEFT : A_UP;           /* :201-209 */
    }
  }
  int q = (int)qLen, t = k - (int)(qLen - tLen);                      /* band column of the corner :247-248 */
  int gmin = S[KB(q, t)], minCol = gmin, minRow = gmin, minColIdx = 0, minRowIdx = 0;
  if (at == ORC_QUERYFIT || at == ORC_FIT) {                           /* :255-279 */
    int set = 0, q2 = (int)qLen;
    for (int t2 = q - k; t2 < q2 + k + 1; t2++) {
      if (t2 < 1 || (uint32_t)t2 > tLen) continue;
      int v = S[KB(q2, k + t2 - q)];
      if (!set || v < minRow) { set = 1; minRow = v; minRowIdx = t2; }
    }
    if (set) { t = k - (q - minRowIdx); q = q2; }
  }
  if (at == ORC_TARGETFIT || at == ORC_FIT) {                          /* :280-304 */
    int set = 0, t2 = k - (int)(qLen - tLen);
    for (int q2 = (int)qLen; q2 >= (int)tLen - k && q2 > 0; q2--) {
      int v = S[KB(q2, k + (int)tLen - q2)];
      if (!set || v < minCol) { minCol = v; set = 1; minColIdx = q2; }
    }
    /* the start column stays the corner's band column even when another row wins (kept as is) */
    if (at == ORC_FIT) { if (minCol < minRow) { t = t2; q = minColIdx; } }
    else { t = t2; q = minColIdx; }
  }
  int opt = S[KB(q, t)];
  path_t p = {0, 0, 0};
  if (at == ORC_GLOBAL || at == ORC_QUERYFIT || at == ORC_FIT || at == ORC_TARGETFIT) {  /* :327-383 */
    for (;;) {
      if (!(q > 0)) break;
      if (at == ORC_TARGETFIT) { if (q < k && k - q == t) break; }
      else {
        if (t < k && k - t == q) break;
        if (at == ORC_FIT && q <= k && k - q == t) break;
      }
      uint8_t a = P[KB(q, t)];
      if (a == A_NONE) break;
      path_push(&p, a);
      if (a == A_DIAG) q--; else if (a == A_UP) { q--; t++; } else if (a == A_LEFT) t--;
    }
  }
  aln->qPos = (uint32_t)q;
  aln->tPos = t < k ? (uint32_t)((k - t) - q) : (uint32_t)((t - k) - q);  /* :394-399 */
  path_reverse(&p);
  arrows_to_alignment(aln, p.a, p.n);
  free(p.a); free(S); free(P);
#undef KB
  res->score = opt; res->alnScore = 0;                                /* alignment.score is never set */
  return 0;
}

/* ---- AffineKBandAlign, AffineKBandAlign.h:12-401 -------------------------------- */
/* Three (qLen+1) x (2k+1) matrices: main S, homopolymer-insertion H, insertion I (:82-87 zero / NoArrow fill).
 * Deletions are linear (bndDel), insertions affine with a second, cheaper state that may only be extended
 * while the query repeats its previous base (:180-188, raw bytes).  Global and QueryFit only: the TargetFit end
 * search reads mirrored band columns (:322-330) and can start the traceback on a never-written NoArrow cell,
 * where the traceback loop (:342-390) does not terminate; other types leave q,t past the matrix. */
#define AK_INF (INT_MAX - 1000)                                        /* :30 */
enum { AK_NONE = 0, AK_DIAG, AK_LEFT, AK_ICLOSE, AK_HCLOSE, AK_OPEN, AK_UP };
static int affine_kband_align(const orc_scorefn *fn, const orc_job *j, orc_result *res, aln_t *aln) {
  int k = j->band, at = j->alignType;
  const int hpO = j->hpInsOpen, hpE = j->hpInsExtend, inO = j->insOpen, inE = j->insExtend, del = j->bndDel;
  if (k < 0 || (at != ORC_GLOBAL && at != ORC_QUERYFIT)) { res->status = ORC_BAD_INPUT; return 0; }
  if (iabs(hpO) >= 1000 || iabs(hpE) >= 1000 || iabs(inO) >= 1000 || iabs(inE) >= 1000 || iabs(del) >= 1000) {
    res->status = ORC_BAD_INPUT; return 0;                             /* INF_SCORE + cost would overflow int */
  }
  uint32_t tLen, qLen; kbounded(j->tLen, j->qLen, (uint32_t)k, &tLen, &qLen);   /* :43 */
  if (at == ORC_QUERYFIT && (tLen == 0 || qLen == 0)) { res->status = ORC_BAD_INPUT; return 0; } /* end search reads unwritten cells */
  int64_t nCols = 2 * (int64_t)k + 1, total = ((int64_t)qLen + 1) * nCols;
  if (total > INT_MAX) { res->status = ORC_BAD_INPUT; return 0; }
  int *S = (int *)calloc((size_t)total, sizeof(int)), *H = (int *)calloc((size_t)total, sizeof(int)), *I = (int *)calloc((size_t)total, sizeof(int));
  uint8_t *PS = (uint8_t *)calloc((size_t)total, 1), *PH = (uint8_t *)calloc((size_t)total, 1), *PI = (uint8_t *)calloc((size_t)total, 1);
#define KB(q_, c_) ((int64_t)(q_) * nCols + (c_))
  I[KB(0, k)] = 0; PI[KB(0, k)] = AK_OPEN;                             /* :93-94 */
  H[KB(0, k)] = 0; PH[KB(0, k)] = AK_OPEN;                             /* :119-120 */
  for (int q = 1; q <= k && q < (int64_t)qLen + 1; q++) {
    I[KB(q, k - q)] = q * inE + inO; PI[KB(q, k - q)] = AK_UP;         /* :96-99 */
    H[KB(q, k - q)] = q * hpE + hpO; PH[KB(q, k - q)] = AK_UP;         /* :122-125 */
  }
  for (int t = k + 1; t < nCols; t++) { H[KB(0, t)] = AK_INF; I[KB(0, t)] = AK_INF; }   /* :127-132 */
  for (int q = 1; q <= k && q < (int64_t)qLen + 1; q++) { S[KB(q, k - q)] = I[KB(q, k - q)]; PS[KB(q, k - q)] = AK_ICLOSE; } /* :134-137 */
  for (int t = 1; t <= k; t++) { S[KB(0, t + k)] = t * del; PS[KB(0, t + k)] = AK_LEFT; }  /* :138-141 (no t < tLen guard) */
  for (int q = 1; q <= (int)qLen; q++) {
    for (int t = q - k; t < q + k + 1; t++) {
      if (t < 1) continue;
      if ((uint32_t)t > tLen) break;                                   /* :159-164 */
      int64_t upper = KB(q - 1, k + t - q + 1), cur = KB(q, k + t - q);
      int inBand = t < q + k;
      int hOpen = inBand ? wadd(S[upper], hpO) : AK_INF;               /* :169-172 */
      int hExt = (q > 1 && j->q[q - 1] == j->q[q - 2] && inBand) ? wadd(H[upper], hpE) : AK_INF;   /* :180-188 */
      int minH;
      if (hOpen < hExt) { PH[cur] = AK_OPEN; minH = hOpen; } else { PH[cur] = AK_UP; minH = hExt; }   /* :195-203 */
      H[cur] = minH;
      int iOpen = inBand ? wadd(S[upper], inO) : AK_INF, iExt = inBand ? wadd(I[upper], inE) : AK_INF;  /* :204-211 */
      int minI;
      if (iOpen < iExt) { PI[cur] = AK_OPEN; minI = iOpen; } else { PI[cur] = AK_UP; minI = iExt; }   /* :213-221 */
      I[cur] = minI;
      int ds = (t == q - k) ? AK_INF : wadd(S[KB(q, k + t - q - 1)], del);                          /* :226-236 */
      int ms = wadd(S[KB(q - 1, k + t - q)], fn->M[g_code[j->q[q - 1]] * 5 + g_code[j->t[t - 1]]]); /* :248 row = query */
      int best = imin(ms, imin(ds, imin(minI, minH)));
      S[cur] = best;
      PS[cur] = best == ms ? AK_DIAG : best == ds ? AK_LEFT : best == minI ? AK_ICLOSE : AK_HCLOSE;   /* :254-269 */
    }
  }
  int q = (int)qLen, t = k - ((int)qLen - (int)tLen);                  /* Global corner :292-295 */
  if (at == ORC_QUERYFIT) {                                            /* :296-311 */
    int best = q - k > 1 ? q - k : 1, minScore = S[KB(qLen, k + best - q)];
    for (int t2 = q - k; t2 < q + k + 1; t2++) {
      if (t2 < 1) continue;
      if ((uint32_t)t2 > tLen) break;
      if (S[KB(qLen, k + t2 - q)] < minScore) { best = t2; minScore = S[KB(qLen, k + t2 - q)]; }
    }
    t = k - ((int)qLen - best);
  }
  int opt = S[KB(q, t)];
  path_t p = {0, 0, 0};
  int mat = 0;                                                         /* 0 Match, 1 AffineHPIns, 2 AffineIns */
  int64_t guard = 4 * total + 16;
  while (q > 0 || (q == 0 && t > k)) {                                 /* :342-390 */
    if (t < 0 || t >= nCols || --guard < 0) { res->status = ORC_PATH_AWRY; break; }
    if (mat == 0) {
      uint8_t a = PS[KB(q, t)];
      if (a == AK_DIAG) { path_push(&p, A_DIAG); q--; }
      else if (a == AK_LEFT) { path_push(&p, A_LEFT); t--; }
      else if (a == AK_ICLOSE) mat = 2;                                /* closes change state only */
      else if (a == AK_HCLOSE) mat = 1;
      else { res->status = ORC_PATH_AWRY; break; }                     /* reference: spins forever on NoArrow */
    } else {
      uint8_t a = mat == 1 ? PH[KB(q, t)] : PI[KB(q, t)];
      if (a == AK_OPEN) mat = 0;
      else if (a != AK_UP) { res->status = ORC_PATH_AWRY; break; }     /* reference: assert(0) */
      path_push(&p, A_UP); q--; t++;                                   /* every step inside an affine matrix emits Up :363-390 */
    }
  }
  if (res->status == ORC_OK) {
    path_reverse(&p);
    arrows_to_alignment(aln, p.a, p.n);                                /* qPos / tPos / nCells / score stay untouched */
  }
  free(p.a); free(S); free(H); free(I); free(PS); free(PH); free(PI);
#undef KB
  res->score = opt; res->alnScore = 0;
  return 0;
}

/* ---- SWAlign, SWAlign.h:18-389 -------------------------------------------------- */
static int sw_align(const orc_scorefn *fn, const orc_job *j, orc_result *res, aln_t *aln) {
  int at = j->alignType;
  int64_t nRows = (int64_t)j->qLen + 1, nCols = (int64_t)j->tLen + 1;
  if (at < 0 || at == ORC_FIT || at > ORC_TPREFIXQSUFFIX || nRows * nCols > INT_MAX) { res->status = ORC_BAD_INPUT; return 0; }
  int *S = (int *)calloc((size_t)(nRows * nCols), sizeof(int));
  uint8_t *P = (uint8_t *)malloc((size_t)(nRows * nCols)); memset(P, A_NONE, (size_t)(nRows * nCols));
#define SW(r_, c_) ((int64_t)(r_) * nCols + (c_))
  int localFam = (at == ORC_LOCAL || at == ORC_ENDANCHORED);
  /* boundaries :49-138 : (row-0 cost per column, row-0 arrow, col-0 cost per row, col-0 arrow) */
  int r0 = 0, c0 = 0; uint8_t r0a = A_LEFT, c0a = A_UP;
  switch (at) {
    case ORC_GLOBAL: case ORC_FRONTANCHORED: r0 = fn->del; c0 = fn->ins; break;
    case ORC_LOCAL: case ORC_ENDANCHORED: r0a = c0a = A_NONE; break;
    case ORC_QUERYFIT: case ORC_OVERLAP: case ORC_TSUFFIXQPREFIX: c0 = fn->ins; break;
    case ORC_TARGETFIT: case ORC_TPREFIXQSUFFIX: r0 = fn->del; break;
  }
  for (int64_t c = 0; c < nCols; c++) { S[SW(0, c)] = (int)(r0 * c); P[SW(0, c)] = r0a; }
  for (int64_t r = 0; r < nRows; r++) { S[SW(r, 0)] = (int)(c0 * r); P[SW(r, 0)] = c0a; }
  P[0] = A_DIAG;                                                      /* :140 */
  int lmin = 0, lminRow = 0, lminCol = 0;
  for (int r = 0; r < (int)j->qLen; r++) {
    for (int c = 0; c < (int)j->tLen; c++) {
      int ms = wadd(match_cost(fn, j, (uint32_t)c, (uint32_t)r), S[SW(r, c)]);
      int qg = wadd(S[SW(r, c + 1)], fn->ins);                         /* :166 (position args irrelevant) */
      int tg = wadd(S[SW(r + 1, c)], fn->del);                         /* :167 */
      int best = imin(ms, imin(qg, tg));
      if (best < lmin) { lmin = best; lminRow = r; lminCol = c; }      /* 0-based loop indices :169-173 */
      if (best > 0 && localFam) { S[SW(r + 1, c + 1)] = 0; P[SW(r + 1, c + 1)] = A_NONE; }
      else { S[SW(r + 1, c + 1)] = best; P[SW(r + 1, c + 1)] = best == ms ? A_DIAG : best == qg ? A_UP : A_LEFT; } /* Diagonal > Up > Left :196-207 */
    }
  }
  int r = 0, c = 0, minRow = 0, minCol = 0;
  if (at == ORC_GLOBAL || at == ORC_ENDANCHORED) { r = minRow = (int)j->qLen; c = minCol = (int)j->tLen; }
  else if (at == ORC_LOCAL || at == ORC_FRONTANCHORED) { r = minRow = lminRow; c = minCol = lminCol; }
  else if (at == ORC_QUERYFIT || at == ORC_OVERLAP || at == ORC_TPREFIXQSUFFIX) {  /* :248-264,:302-314 */
    if (nCols < 2) { res->status = ORC_BAD_INPUT; goto out; }
    int best = S[SW(nRows - 1, 1)]; minCol = 1;
    for (int cc = 2; cc < nCols; cc++) if (S[SW(nRows - 1, cc)] < best) { best = S[SW(nRows - 1, cc)]; minCol = cc; }
    c = minCol; r = minRow = (int)nRows - 1;
  } else { /* TargetFit :265-284 (minRow uninitialised in the reference when row 1 wins), TSuffixQPrefix :285-301 */
    if (nRows < 2) { res->status = ORC_BAD_INPUT; goto out; }
    int best = S[SW(1, nCols - 1)]; minRow = 1;
    for (int rr = 2; rr < nRows; rr++) if (S[SW(rr, nCols - 1)] < best) { best = S[SW(rr, nCols - 1)]; minRow = rr; }
    r = minRow; c = minCol = (int)nCols - 1;
  }
  {
    path_t p = {0, 0, 0};
    for (;;) {                                                        /* :324-353 */
      int go;
      if (at == ORC_GLOBAL || at == ORC_FRONTANCHORED) go = (r > 0 || c > 0);
      else if (at == ORC_QUERYFIT || at == ORC_OVERLAP || at == ORC_TSUFFIXQPREFIX) go = r > 0;
      else if (at == ORC_TPREFIXQSUFFIX || at == ORC_TARGETFIT) go = c > 0;
      else go = (r > 0 && c > 0 && P[SW(r, c)] != A_NONE);
      if (!go) break;
      uint8_t a = P[SW(r, c)];
      path_push(&p, a);
      if (a == A_DIAG) { r--; c--; } else if (a == A_UP) r--; else if (a == A_LEFT) c--;
      else { res->status = ORC_PATH_AWRY; break; } /* reference would spin forever */
    }
    path_reverse(&p);
    if (p.n > 0) arrows_to_alignment(aln, p.a, p.n);
    free(p.a);
    if (at != ORC_GLOBAL && at != ORC_FRONTANCHORED && at != ORC_OVERLAP) { aln->qPos = (uint32_t)r; aln->tPos = (uint32_t)c; } /* :367-380 */
  }
  res->score = S[SW(minRow, minCol)];                                  /* :388 */
  res->alnScore = 0; res->nCells = 0;
out:
  free(S); free(P);
#undef SW
  return 0;
}

/* ---- ComputeAlignmentStats, AlignmentUtils.h:535-584 over CreateAlignmentStrings :390-533
 *      and ComputeAlignmentScore(string,...) :60-124; evaluated column by column without
 *      materialising the three strings. ------------------------------------------------ */
typedef struct {
  const orc_scorefn *fn; int affine;
  int nMatch, nMismatch, nIns, nDel; int64_t len; int score;
  int runLen, runLastIsDel; /* open gap run for affine scoring */
} stat_t;
static void stat_close_run(stat_t *s) {
  if (s->runLen) {
    int aff = s->runLen * s->fn->affineExtend + s->fn->affineOpen;                 /* :87 */
    int lin = (s->runLastIsDel ? s->fn->del : s->fn->ins) * s->runLen;              /* :89-94 typed by last column */
    s->score += lin < aff ? lin : aff;
    s->runLen = 0;
  }
}
static void stat_pair(stat_t *s, uint8_t qc, uint8_t tc) {
  stat_close_run(s);
  if (g_code[tc] == g_code[qc]) s->nMatch++; else s->nMismatch++;
  s->score += s->fn->M[(g_code[qc] % 5) * 5 + (g_code[tc] % 5)];
  s->len++;
}
static void stat_gap(stat_t *s, int isDel, int n) {
  if (isDel) s->nDel += n; else s->nIns += n;
  s->len += n;
  if (s->affine) { s->runLen += n; s->runLastIsDel = isDel; }
  else s->score += n * (isDel ? s->fn->del : s->fn->ins);
}
static void compute_stats(const orc_scorefn *fn, const orc_job *j, const aln_t *a, int affine, orc_result *res) {
  stat_t s; memset(&s, 0, sizeof s); s.fn = fn; s.affine = affine;
  uint32_t q = a->qPos, t = a->tPos;
  if (a->nBlocks) {
    uint32_t g = 0;
    if (a->nGapLists == 0) { /* CreateAlignmentStrings :409-450: leading offset of block 0 */
      uint32_t qp = a->blocks[0], tp = a->blocks[1], common = qp < tp ? qp : tp;
      for (uint32_t i = 0; i < common; i++) stat_pair(&s, j->q[q++], j->t[t++]);
      if (tp - common) { stat_gap(&s, 1, (int)(tp - common)); t += tp - common; }
      if (qp - common) { stat_gap(&s, 0, (int)(qp - common)); q += qp - common; }
    } else {
      for (uint32_t i = 0; i < a->gapCounts[0]; i++, g++) {
        int isDel = a->gaps[2 * g] == 0, n = a->gaps[2 * g + 1];
        stat_gap(&s, isDel, n); if (isDel) t += (uint32_t)n; else q += (uint32_t)n;
      }
    }
    for (uint32_t b = 0; b < a->nBlocks; b++) {
      for (uint32_t l = 0; l < a->blocks[3 * b + 2]; l++) stat_pair(&s, j->q[q++], j->t[t++]);
      if (b + 1 == a->nBlocks) continue;
      if (a->nGapLists > 0) {
        for (uint32_t i = 0; i < a->gapCounts[b + 1]; i++, g++) {
          int isDel = a->gaps[2 * g] == 0, n = a->gaps[2 * g + 1];
          stat_gap(&s, isDel, n); if (isDel) t += (uint32_t)n; else q += (uint32_t)n;
        }
      } else { /* :497-529 */
        int qg = (int)(a->blocks[3 * (b + 1)] - a->blocks[3 * b] - a->blocks[3 * b + 2]);
        int tg = (int)(a->blocks[3 * (b + 1) + 1] - a->blocks[3 * b + 1] - a->blocks[3 * b + 2]);
        if (qg > 0 || tg > 0) {
          int common = qg > tg ? tg : qg; tg -= common; qg -= common;
          if (qg > 0) { stat_gap(&s, 0, qg); q += (uint32_t)qg; }
          if (tg > 0) { stat_gap(&s, 1, tg); t += (uint32_t)tg; }
          for (int i = 0; i < common; i++) stat_pair(&s, j->q[q++], j->t[t++]);
        }
      }
    }
    stat_close_run(&s);
  }
  res->nMatch = s.nMatch; res->nMismatch = s.nMismatch; res->nIns = s.nIns; res->nDel = s.nDel;
  res->statsScore = s.score;
  /* :566-576: tp+qp>0 and both strings have the same length s.len */
  res->pctSimilarity = s.len > 0 ? (float)((s.nMatch * 2.0) / (double)(2 * s.len) * 100) : 0.0f;
}

/* ---- entry points -------------------------------------------------------------- */
static int check_bases(const uint8_t *s, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) if (g_code[s[i]] > 4) return 0;
  return 1;
}

int orc_align(const orc_scorefn *fn, const orc_job *job, orc_result *res, uint32_t *blocks, uint32_t capBlocks,
              uint32_t *gapCounts, uint32_t capGapLists, int32_t *gaps, uint32_t capGaps) {
  pthread_once(&g_once, init_codes);
  memset(res, 0, sizeof *res);
  aln_t a; memset(&a, 0, sizeof a);
  a.blocks = blocks; a.capBlocks = capBlocks; a.gapCounts = gapCounts; a.capGapLists = capGapLists;
  a.gaps = gaps; a.capGaps = capGaps;
  if (!check_bases(job->q, job->qLen) || !check_bases(job->t, job->tLen) ||
      (fn->kind == ORC_FN_QUALITY && !job->qual) ||
      (fn->kind == ORC_FN_IDS && (!job->insQV || !job->subQV || !job->subTag || job->algo == ORC_SW))) {
    /* SWAlign x IDS: the reference passes transposed positions (SWAlign.h:166-167) and reads out of bounds */
    res->status = ORC_BAD_INPUT; return 0;
  }
  switch (job->algo) {
    case ORC_GUIDED: guided_align(fn, job, 0, res, &a); break;
    case ORC_AFFINE_GUIDED: guided_align(fn, job, 1, res, &a); break;
    case ORC_KBAND: kband_align(fn, job, res, &a); break;
    case ORC_SW: sw_align(fn, job, res, &a); break;
    case ORC_AFFINE_KBAND: affine_kband_align(fn, job, res, &a); break;
    default: res->status = ORC_BAD_INPUT;
  }
  res->qPos = a.qPos; res->tPos = a.tPos;
  res->nBlocks = a.nBlocks; res->nGapLists = a.nGapLists; res->nGaps = a.nGaps;
  if (a.overflow) return ORC_OVERFLOW;
  if (job->doStats && res->status <= ORC_EMPTY_GUIDE) compute_stats(fn, job, &a, job->statsAffine, res);
  return 0;
}

typedef struct { const orc_scorefn *fn; const orc_job *jobs; uint32_t n; uint32_t *next; pthread_mutex_t *mu; int64_t cells, sum; } rp_t;
static void *replay_worker(void *arg) {
  rp_t *w = (rp_t *)arg;
  uint32_t capB = 1 << 16, capG = 1 << 17;
  uint32_t *blocks = (uint32_t *)malloc(sizeof(uint32_t) * 3 * capB), *gc = (uint32_t *)malloc(sizeof(uint32_t) * (capB + 1));
  int32_t *gaps = (int32_t *)malloc(sizeof(int32_t) * 2 * capG);
  for (;;) {
    pthread_mutex_lock(w->mu); uint32_t i = (*w->next)++; pthread_mutex_unlock(w->mu);
    if (i >= w->n) break;
    orc_result r;
    orc_align(w->fn, &w->jobs[i], &r, blocks, capB, gc, capB + 1, gaps, capG);
    w->cells += r.nCells; w->sum += r.score + (int64_t)r.nBlocks;
  }
  free(blocks); free(gc); free(gaps);
  return NULL;
}
int64_t orc_replay(const orc_scorefn *fn, const orc_job *jobs, uint32_t n, int nThreads, int64_t *scoreSum) {
  if (nThreads < 1) nThreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
  rp_t *w = (rp_t *)calloc((size_t)nThreads, sizeof(rp_t));
  uint32_t next = 0; pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  for (int i = 0; i < nThreads; i++) { w[i].fn = fn; w[i].jobs = jobs; w[i].n = n; w[i].next = &next; w[i].mu = &mu; pthread_create(&th[i], NULL, replay_worker, &w[i]); }
  int64_t cells = 0, sum = 0;
  for (int i = 0; i < nThreads; i++) { pthread_join(th[i], NULL); cells += w[i].cells; sum += w[i].sum; }
  if (scoreSum) *scoreSum = sum;
  free(th); free(w);
  return cells;
}


/* ------------------------------------------------------------------------------------------------------------------
 * SAM CIGAR core, restated from printers/SAMPrinter.h:203-293 for alignments whose gap lists are filled (the
 * `nGaps > 0` branch): AddGaps(gaps[0]) (:120-137: Gap::Query -> 'D' and tPos += length, Gap::Target -> 'I' and
 * qPos += length), then per block AddUngappedOperations (:138-166: alternating maximal runs of unequal 'X' and equal '='
 * RAW bytes of the two aligned sequences, qPos / tPos being running counters, not the block's own coordinates) and
 * AddGaps(gaps[b+1]).  q / t are the sequences the candidate's qAlignedSeq / tAlignedSeq reference. */
static int cig_push(uint32_t *ops, uint32_t cap, uint32_t *n, uint32_t len, uint32_t code) {
  if (*n >= cap) return -1;
  ops[(*n)++] = (len << 4) | code;
  return 0;
}
int orc_cigar_from(const uint8_t *q, const uint8_t *t, uint32_t qPos, uint32_t tPos,
                   const uint32_t *blocks, uint32_t nBlocks, const uint32_t *gapCounts, uint32_t nGapLists,
                   const int32_t *gaps, uint32_t *ops, uint32_t capOps) {
  uint32_t n = 0, g = 0, b, i, j;
  if (nBlocks == 0) return 0;
  qPos += blocks[0]; tPos += blocks[1];                     /* CreateCIGARString :362-363 */
  for (b = 0; b <= nBlocks; b++) {
    if (b > 0) {                                            /* AddUngappedOperations(b-1) */
      const uint32_t len = blocks[3 * (b - 1) + 2];
      i = 0;
      while (i < len) {
        uint32_t s0 = i;
        while (i < len && q[qPos + i] != t[tPos + i]) i++;
        if (i > s0 && cig_push(ops, capOps, &n, i - s0, 8)) return -1;
        s0 = i;
        while (i < len && q[qPos + i] == t[tPos + i]) i++;
        if (i > s0 && cig_push(ops, capOps, &n, i - s0, 7)) return -1;
      }
      qPos += len; tPos += len;
    }
    if (b < nGapLists) {                                    /* AddGaps(b) */
      for (j = 0; j < gapCounts[b]; j++, g++) {
        const int32_t seq = gaps[2 * g], len = gaps[2 * g + 1];
        if (seq == 0) { if (cig_push(ops, capOps, &n, (uint32_t)len, 2)) return -1; tPos += (uint32_t)len; }
        else if (seq == 1) { if (cig_push(ops, capOps, &n, (uint32_t)len, 1)) return -1; qPos += (uint32_t)len; }
      }
    }
  }
  return (int)n;
}

/* ------------------------------------------------------------------------------------------------------------------
 * The three strings of the m5 / stick printers, restated from CreateAlignmentStrings (AlignmentUtils.h:390-533) and
 * AppendGapCharacters (:366-387).  q / t are the sequences the alignment indexes (positions start at qPos / tPos).
 * With gap lists: gaps[0] columns carry '*' in the middle string, block columns '|' or '*' by TwoBit equality (every
 * non-ACGT byte maps to 255 there, so N pairs with any other ambiguity code as a match), gaps[b+1] columns ' '.
 * Without gap lists (SDPAlign output): leading offsets of block 0 and the gaps between blocks are laid out as
 * :409-447 / :497-529 do.  out* need room for capOut bytes each; returns the common length, -1 on overflow. */
static int two_bit(uint8_t c) {
  if (c <= 7) return c & 3;
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 255; }
}
#define STR_PUSH(tc, ac, qc) do { if (n >= capOut) return -1; textStr[n] = (char)(tc); alignStr[n] = (char)(ac); queryStr[n] = (char)(qc); n++; } while (0)
int orc_alignment_strings(const uint8_t *query, const uint8_t *text, uint32_t qPos, uint32_t tPos,
                          const uint32_t *blocks, uint32_t nBlocks, const uint32_t *gapCounts, uint32_t nGapLists,
                          const int32_t *gaps, char *textStr, char *alignStr, char *queryStr, uint32_t capOut) {
  uint32_t q = qPos, t = tPos, n = 0, b, bl, g = 0, gi, p;
  if (nBlocks == 0) return 0;
  if (nGapLists == 0) {                                     /* :409-447 */
    if (blocks[0] > 0 || blocks[1] > 0) {
      uint32_t qp = blocks[0], tp = blocks[1];
      int common = (int)qp;
      if ((uint32_t)common > tp) common = (int)tp;
      for (p = 0; p < (uint32_t)common; p++) { STR_PUSH(text[t], '*', query[q]); t++; q++; }
      tp -= (uint32_t)common; qp -= (uint32_t)common;
      for (p = 0; p < tp; p++) { STR_PUSH(text[t], ' ', '-'); t++; }
      for (p = 0; p < qp; p++) { STR_PUSH('-', ' ', query[q]); q++; }
    }
  } else {                                                  /* :453-460: the list before the first block */
    for (gi = 0; gi < gapCounts[0]; gi++, g++) {
      int k;
      for (k = 0; k < gaps[2 * g + 1]; k++) {
        if (gaps[2 * g] == 0) { STR_PUSH(text[t], '*', '-'); t++; }
        else if (gaps[2 * g] == 1) { STR_PUSH('-', '*', query[q]); q++; }
      }
    }
  }
  for (b = 0; b < nBlocks; b++) {
    for (bl = 0; bl < blocks[3 * b + 2]; bl++) {
      STR_PUSH(text[t], two_bit(query[q]) != two_bit(text[t]) ? '*' : '|', query[q]);
      q++; t++;
    }
    if (b == nBlocks - 1) continue;
    if (nGapLists > 0) {
      for (gi = 0; gi < gapCounts[b + 1]; gi++, g++) {
        int k;
        for (k = 0; k < gaps[2 * g + 1]; k++) {
          if (gaps[2 * g] == 0) { STR_PUSH(text[t], ' ', '-'); t++; }
          else if (gaps[2 * g] == 1) { STR_PUSH('-', ' ', query[q]); q++; }
        }
      }
    } else {                                                /* :497-529 */
      int queryGapLen = (int)(blocks[3 * (b + 1)] - blocks[3 * b] - blocks[3 * b + 2]);
      int textGapLen = (int)(blocks[3 * (b + 1) + 1] - blocks[3 * b + 1] - blocks[3 * b + 2]);
      if (queryGapLen > 0 || textGapLen > 0) {
        int common = queryGapLen, k;
        if (queryGapLen > textGapLen) common = textGapLen;
        textGapLen -= common; queryGapLen -= common;
        for (k = 0; k < queryGapLen; k++, q++) STR_PUSH('-', ' ', query[q]);
        for (k = 0; k < textGapLen; k++, t++) STR_PUSH(text[t], ' ', '-');
        for (k = 0; k < common; k++) { STR_PUSH(text[t], ' ', query[q]); t++; q++; }
      }
    }
  }
  return (int)n;
}
