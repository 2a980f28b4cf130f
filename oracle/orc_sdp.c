/*
 * oracle/orc_sdp.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the chaining step of SDPAlign, the next row of the hot-path scope table:
 *   SDPLongestCommonSubsequence   common/algorithms/alignment/sdp/SparseDynamicProgramming.h:71-322
 *   StoreAbove                    :51-69        IndelPenalty  :27-49
 *   SDPSet (Predecessor / Successor / Insert / Delete / Member)   sdp/SDPSet.h:16-120
 *   Fragment ordering             sdp/SDPFragment.h:62-93, sdp/FragmentSort.h, sdp/SDPColumn.h
 * It takes a fragment set with UNIQUE (x, y) -- what SDPAlign.h:249-262 hands over after its sort + de-duplication --
 * so every sort below has a unique key and the result does not depend on the sort implementation.
 * Quirks of the reference that change results are kept and marked "as the reference".
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "orc_align.h"

typedef struct { uint32_t x, y, weight, length; int index, chainPrev, cost, above; uint32_t chainLength; } Frag;
typedef struct { int col, opt; } Col;                       /* SDPColumn */
typedef struct { uint32_t x, y; int cost, index; } Swept;   /* the copy a Fragment leaves in the sweep set */

static int cmp_xy(const void *a, const void *b) {           /* LexicographicFragmentSort: LessThanXY */
  const Frag *p = (const Frag *)a, *q = (const Frag *)b;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  return p->y < q->y ? -1 : (p->y > q->y ? 1 : 0);
}
static int cmp_yx(const void *a, const void *b) {           /* LexicographicFragmentSortByY: LessThanYX */
  const Frag *p = (const Frag *)a, *q = (const Frag *)b;
  if (p->y != q->y) return p->y < q->y ? -1 : 1;
  return p->x < q->x ? -1 : (p->x > q->x ? 1 : 0);
}
/* Fragment::operator< : by diagonal (int)(y - x), then by x (SDPFragment.h:78-93) */
static int swept_less(uint32_t ax, uint32_t ay, uint32_t bx, uint32_t by) {
  const int da = (int)(ay - ax), db = (int)by - (int)bx;
  if (da < db) return 1;
  if (da == db) return ax < bx;
  return 0;
}
static int indel_penalty(int x1, int y1, int x2, int y2, int insertion, int deletion) {   /* :27-49 */
  const int drift = (x1 - y1) - (x2 - y2);
  if (drift > 0) return (int)((1.0 * drift) * insertion);
  if (drift < 0) return (int)((-1.0 * drift) * deletion);
  return 0;
}

/* first index whose key is not less than (x, y) in the sweep set (std::set::lower_bound) */
static int swept_lower_bound(const Swept *s, int n, uint32_t x, uint32_t y) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (swept_less(s[mid].x, s[mid].y, x, y)) lo = mid + 1; else hi = mid; }
  return lo;
}
static int col_lower_bound(const Col *c, int n, int col) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid].col < col) lo = mid + 1; else hi = mid; }
  return lo;
}

int orc_sdp_chain(const uint32_t *frags, uint32_t n, uint32_t queryLength, uint32_t fragmentLength,
                  int insertion, int deletion, int match, int alignType, int32_t *chain, uint32_t capChain) {
  if (n < 1) return 0;
  Frag *f = (Frag *)malloc(sizeof(Frag) * n);
  Col *cols = (Col *)malloc(sizeof(Col) * (n + 1));
  Swept *sw = (Swept *)malloc(sizeof(Swept) * (n + 1));
  int nCols = 0, nSw = 0;
  uint32_t i;
  for (i = 0; i < n; i++) {                                   /* Fragment(x, y, weight), length set by SDPAlign.h:204-215 */
    f[i].x = frags[4 * i]; f[i].y = frags[4 * i + 1]; f[i].length = frags[4 * i + 2]; f[i].weight = frags[4 * i + 3];
    f[i].chainPrev = 0; f[i].cost = 0; f[i].above = -1; f[i].index = 0; f[i].chainLength = 0;
  }
  qsort(f, n, sizeof(Frag), cmp_xy);                          /* :80 */
  for (i = 0; i < n; i++) f[i].index = (int)i;               /* :89-91 */
  /* StoreAbove :51-69: neighbours in (y, x) order; the earlier one is "above" when it still covers this x */
  qsort(f, n, sizeof(Frag), cmp_yx);
  for (i = 1; i < n; i++)
    if (f[i - 1].x <= f[i].x && f[i - 1].x + f[i - 1].length > f[i].x && f[i - 1].y < f[i].y) f[i].above = f[i - 1].index;
  qsort(f, n, sizeof(Frag), cmp_xy);

  uint32_t sweepRow = f[0].x, fSweep = 0, fTrail = 0, maxChainLength = 0;
  int maxChainFragment = -1, minFragmentCost = INT_MAX, minFragmentIndex = -1;
  for (; sweepRow < queryLength + fragmentLength; sweepRow++) {                    /* :108 */
    const uint32_t startF = fSweep;
    while (fSweep < n && f[fSweep].x == sweepRow) {
      Frag *c = &f[fSweep];
      int cp = INT_MAX, cl = INT_MAX, ca = INT_MAX, foundPrev = 0, predOpt = -1, predIndex = -1;
      /* colSet.Predecessor: the column with the greatest col <= y (SDPSet.h:95-120) */
      if (nCols > 0) {
        int it = col_lower_bound(cols, nCols, (int)c->y);
        int have = 0;
        if (it < nCols && cols[it].col == (int)c->y) have = 1;
        else { if (it != 0) --it; if (!((int)c->y < cols[it].col)) have = 1; }
        if (have) {
          predOpt = cols[it].opt;
          const int dist = abs((int)(c->x + c->y) - (int)(f[predOpt].x + f[predOpt].y));
          cp = (int)((uint32_t)f[predOpt].cost + (uint32_t)(int)sqrt((double)dist) - c->length);   /* :133-136 */
          foundPrev = 1;
        }
      }
      /* sweepSet.Predecessor: the swept fragment with the greatest (diagonal, x) <= this one's */
      if (nSw > 0) {
        int it = swept_lower_bound(sw, nSw, c->x, c->y);
        int have = 0;
        if (it < nSw && !swept_less(c->x, c->y, sw[it].x, sw[it].y)) have = 1;   /* equivalent element */
        else { if (it != 0) --it; if (!swept_less(c->x, c->y, sw[it].x, sw[it].y)) have = 1; }
        if (have) {
          const Swept *p = &sw[it];
          const int overlap = (int)(fragmentLength - (c->y - p->y)) * match;                 /* :157 */
          cl = p->cost + (overlap < 0 ? overlap : 0) + indel_penalty((int)c->x, (int)c->y, (int)p->x, (int)p->y, insertion, deletion);
          predIndex = p->index;
          foundPrev = 1;
        }
      }
      if (c->above >= 0) {                                                                     /* :164-175 */
        const Frag *a = &f[c->above];
        ca = (int)((uint32_t)a->cost + (fragmentLength - (uint32_t)(int)(c->y - a->y)) * (uint32_t)match +
                   (uint32_t)indel_penalty((int)c->x, (int)c->y, (int)a->x, (int)a->y, insertion, deletion));
        foundPrev = 1;
      }
      int minCost = cl < ca ? cl : ca;
      minCost = cp < minCost ? cp : minCost;                                                   /* MIN(cp, MIN(cl, ca)) */
      if (foundPrev && (alignType == ORC_GLOBAL || (alignType == ORC_LOCAL && minCost < 0))) {
        c->cost = (int)((uint32_t)minCost - c->weight);
        if (minCost == cp) c->chainPrev = predOpt;
        else if (minCost == cl) c->chainPrev = predIndex;
        else if (minCost == ca) c->chainPrev = c->above;
        c->chainLength = f[c->chainPrev].chainLength + 1;
      } else if (alignType == ORC_GLOBAL) {
        c->chainPrev = -1;
        c->cost = (int)((c->x + c->y) * (uint32_t)deletion + fragmentLength * (uint32_t)match - c->weight);   /* :211 */
        c->chainLength = 1;
      } else if (alignType == ORC_LOCAL) {
        c->chainPrev = -1;
        c->cost = (int)(fragmentLength * (uint32_t)match - c->weight);
        c->chainLength = 1;
      }
      /* any other alignType: the reference leaves cost / chainPrev / chainLength as constructed (not exercised) */
      if (minFragmentCost > c->cost) { minFragmentCost = c->cost; minFragmentIndex = (int)fSweep; }
      if (c->chainLength > maxChainLength) { maxChainLength = c->chainLength; maxChainFragment = (int)fSweep; }
      fSweep++;
    }
    /* the row's fragments enter the sweep set */
    for (fSweep = startF; fSweep < n && f[fSweep].x == sweepRow; fSweep++) {
      int it = swept_lower_bound(sw, nSw, f[fSweep].x, f[fSweep].y);
      if (it < nSw && !swept_less(f[fSweep].x, f[fSweep].y, sw[it].x, sw[it].y)) {             /* replace an equivalent one */
        sw[it].x = f[fSweep].x; sw[it].y = f[fSweep].y; sw[it].cost = f[fSweep].cost; sw[it].index = f[fSweep].index;
      } else {
        memmove(&sw[it + 1], &sw[it], sizeof(Swept) * (size_t)(nSw - it));
        sw[it].x = f[fSweep].x; sw[it].y = f[fSweep].y; sw[it].cost = f[fSweep].cost; sw[it].index = f[fSweep].index;
        nSw++;
      }
    }
    /* fragments fragmentLength + 1 rows back leave the sweep set and may become their column's representative */
    if (sweepRow >= fragmentLength + 1) {
      const uint32_t trailRow = sweepRow - fragmentLength - 1;
      while (fTrail < n && f[fTrail].x == trailRow) {
        const int y = (int)f[fTrail].y;
        int it = col_lower_bound(cols, nCols, y);
        int storeCol;
        if (it < nCols && cols[it].col == y) storeCol = f[cols[it].opt].cost < f[fTrail].cost;   /* as the reference (:258-262): the
                                                        existing entry is replaced when it is the CHEAPER one */
        else storeCol = 1;
        if (storeCol) {
          if (it < nCols && cols[it].col == y) cols[it].opt = (int)fTrail;
          else { memmove(&cols[it + 1], &cols[it], sizeof(Col) * (size_t)(nCols - it)); cols[it].col = y; cols[it].opt = (int)fTrail; nCols++; }
          /* Successor answers "none" for sets of fewer than two elements (SDPSet.h:79-81) */
          while (nCols >= 2 && it + 1 < nCols && f[cols[it + 1].opt].cost > f[fTrail].cost) {
            memmove(&cols[it + 1], &cols[it + 2], sizeof(Col) * (size_t)(nCols - it - 2));
            nCols--;
          }
        }
        {   /* sweepSet.Delete */
          int s = swept_lower_bound(sw, nSw, f[fTrail].x, f[fTrail].y);
          if (s < nSw && sw[s].x == f[fTrail].x && sw[s].y == f[fTrail].y) { memmove(&sw[s], &sw[s + 1], sizeof(Swept) * (size_t)(nSw - s - 1)); nSw--; }
        }
        ++fTrail;
      }
    }
  }
  if (alignType == ORC_LOCAL) maxChainFragment = minFragmentIndex;
  uint32_t len = 0;
  int k;
  for (k = maxChainFragment; k != -1; k = f[k].chainPrev) len++;
  if (len > capChain) { free(f); free(cols); free(sw); return -1; }
  i = len;
  for (k = maxChainFragment; k != -1; k = f[k].chainPrev) chain[--i] = k;
  free(f); free(cols); free(sw);
  return (int)len;
}
