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

/* ------------------------------------------------------------------------------------------------------------------
 * The fragment set SDPAlign builds in front of the chain (common/algorithms/alignment/SDPAlign.h:133-262):
 *   - the target's first / last sdpPrefixLength bases and the whole target are turned into (tuple, pos) lists
 *     (SequenceToTupleList, tuples/DNATuple.h:309-358: every k-mer of every ACGT-only span; tuples are read right to
 *     left, :55-83) with word sizes small = min(wordSize, SDP_DETAILED_WORD_SIZE = 5), small, wordSize, and sorted by
 *     (tuple, pos) -- a unique key;
 *   - every valid k-mer of the query's first / last sdpPrefixLength bases and of the whole query is looked up
 *     (StoreMatchingPositions, tuples/TupleMatching.h:37-62; FindAll, tuples/TupleList.h:88-101: all positions of an
 *     equal tuple, ascending);
 *   - weights are wordSize everywhere, lengths small / wordSize (:204-215); suffix matches are shifted to absolute
 *     coordinates (:223-230); the three sets are concatenated prefix, middle, suffix (:235-236), sorted by (x, y) with
 *     std::sort and de-duplicated keeping the FIRST of equal (x, y) (:249-262).
 * The same (x, y) regularly arrives from the prefix set (length 5) and from the middle set (length 11): which one
 * survives is decided by the order std::sort leaves equal keys in.  The reference is built with libstdc++ (GCC 13), so
 * that algorithm is restated here -- introsort: median-of-three quicksort down to 16 elements, then insertion sort
 * (bits/stl_algo.h: __introsort_loop, __unguarded_partition_pivot, __move_median_to_first, __final_insertion_sort).
 * including its heapsort escape past a recursion depth of 2 * log2(n), which the nearly sorted concatenation does reach. */
#define ORC_SDP_UNPINNED (-2)

typedef struct { uint32_t x, y, length, weight; } Frag4;
static int f4_less(const Frag4 *a, const Frag4 *b) { return a->x < b->x || (a->x == b->x && a->y < b->y); }   /* LessThanXY */
static void f4_swap(Frag4 *a, Frag4 *b) { const Frag4 t = *a; *a = *b; *b = t; }

static void move_median_to_first(Frag4 *result, Frag4 *a, Frag4 *b, Frag4 *c) {
  if (f4_less(a, b)) {
    if (f4_less(b, c)) f4_swap(result, b);
    else if (f4_less(a, c)) f4_swap(result, c);
    else f4_swap(result, a);
  } else if (f4_less(a, c)) f4_swap(result, a);
  else if (f4_less(b, c)) f4_swap(result, c);
  else f4_swap(result, b);
}
static Frag4 *unguarded_partition(Frag4 *first, Frag4 *last, Frag4 *pivot) {
  for (;;) {
    while (f4_less(first, pivot)) ++first;
    --last;
    while (f4_less(pivot, last)) --last;
    if (!(first < last)) return first;
    f4_swap(first, last);
    ++first;
  }
}
/* the heapsort escape: std::__partial_sort(first, last, last) = __make_heap + __sort_heap (bits/stl_heap.h) */
static void push_heap_(Frag4 *first, long hole, long top, Frag4 value) {
  long parent = (hole - 1) / 2;
  while (hole > top && f4_less(first + parent, &value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
  first[hole] = value;
}
static void adjust_heap(Frag4 *first, long hole, long len, Frag4 value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (f4_less(first + child, first + (child - 1))) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_(first, hole, top, value);
}
static void heap_sort(Frag4 *first, Frag4 *last) {
  const long len = last - first;
  long parent;
  if (len >= 2) for (parent = (len - 2) / 2;; parent--) { adjust_heap(first, parent, len, first[parent]); if (parent == 0) break; }
  while (last - first > 1) { --last; { const Frag4 value = *last; *last = *first; adjust_heap(first, 0, last - first, value); } }
}
static int introsort_loop(Frag4 *first, Frag4 *last, long depth) {
  while (last - first > 16) {
    if (depth == 0) { heap_sort(first, last); return 0; }
    --depth;
    Frag4 *mid = first + (last - first) / 2;
    move_median_to_first(first, first + 1, mid, last - 1);
    Frag4 *cut = unguarded_partition(first + 1, last, first);
    if (introsort_loop(cut, last, depth)) return ORC_SDP_UNPINNED;
    last = cut;
  }
  return 0;
}
static void unguarded_linear_insert(Frag4 *last) {
  const Frag4 val = *last;
  Frag4 *next = last - 1;
  while (f4_less(&val, next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void insertion_sort(Frag4 *first, Frag4 *last) {
  Frag4 *i;
  if (first == last) return;
  for (i = first + 1; i != last; ++i) {
    if (f4_less(i, first)) { const Frag4 val = *i; memmove(first + 1, first, sizeof(Frag4) * (size_t)(i - first)); *first = val; }
    else unguarded_linear_insert(i);
  }
}
static int std_sort_xy(Frag4 *first, Frag4 *last) {
  Frag4 *i;
  long n = last - first, lg = 0;
  if (first == last) return 0;
  while ((1L << (lg + 1)) <= n) lg++;                      /* std::__lg */
  if (introsort_loop(first, last, lg * 2)) return ORC_SDP_UNPINNED;
  if (last - first > 16) { insertion_sort(first, first + 16); for (i = first + 16; i != last; ++i) unguarded_linear_insert(i); }
  else insertion_sort(first, last);
  return 0;
}

static int base2(uint8_t c) {                               /* TwoBit where ThreeBit <= 3 (NucConversion.h:7-84), else -1 */
  switch (c) {
    case 0: case 'A': case 'a': return 0;
    case 1: case 'C': case 'c': return 1;
    case 2: case 'G': case 'g': return 2;
    case 3: case 'T': case 't': return 3;
    default: return -1;
  }
}
typedef struct { uint64_t tuple; uint32_t pos; } TPos;
static int tpos_cmp(const void *a, const void *b) {
  const TPos *p = (const TPos *)a, *q = (const TPos *)b;
  if (p->tuple != q->tuple) return p->tuple < q->tuple ? -1 : 1;
  return p->pos < q->pos ? -1 : (p->pos > q->pos ? 1 : 0);
}
/* every k-mer of every ACGT-only span of s, read right to left (the leftmost base in the lowest two bits) */
static uint32_t tuple_list(const uint8_t *s, uint32_t len, int k, TPos *out) {
  uint32_t n = 0, i, run = 0;
  if (k <= 0 || len < (uint32_t)k) return 0;
  for (i = 0; i < len; i++) {
    run = base2(s[i]) >= 0 ? run + 1 : 0;
    if (run >= (uint32_t)k) {
      uint64_t v = 0; int j;
      for (j = k - 1; j >= 0; j--) v = (v << 2) | (uint64_t)base2(s[i - (uint32_t)(k - 1) + (uint32_t)j]);
      out[n].tuple = v; out[n].pos = i - (uint32_t)(k - 1); n++;
    }
  }
  qsort(out, n, sizeof(TPos), tpos_cmp);
  return n;
}
/* StoreMatchingPositions with maxMatches = 0: (s, pos) for every target position holding the query's k-mer at s */
static uint32_t match_positions(const uint8_t *q, uint32_t qLen, int k, const TPos *list, uint32_t nList,
                                uint32_t xOff, uint32_t yOff, uint32_t length, uint32_t weight, Frag4 *out, uint32_t n, uint32_t cap,
                                int maxMatches) {
  uint32_t s, run = 0;
  if (k <= 0 || qLen < (uint32_t)k) return n;
  for (s = 0; s + (uint32_t)k <= qLen + 0u; s++) {
    /* validity of the window [s, s + k): recomputed per position, which is what the res state machine amounts to */
    uint64_t v = 0; int j, ok = 1;
    for (j = k - 1; j >= 0; j--) { const int b = base2(q[s + (uint32_t)j]); if (b < 0) { ok = 0; break; } v = (v << 2) | (uint64_t)b; }
    (void)run;
    if (!ok) continue;
    uint32_t lo = 0, hi = nList;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (list[mid].tuple < v) lo = mid + 1; else hi = mid; }
    if (maxMatches != 0) {                                   /* positions with more matches than that are skipped (:50) */
      uint32_t e = lo;
      while (e < nList && list[e].tuple == v) e++;
      if ((long)(e - lo) > (long)maxMatches) continue;
    }
    for (; lo < nList && list[lo].tuple == v; lo++) {
      if (n < cap) { out[n].x = s + xOff; out[n].y = list[lo].pos + yOff; out[n].length = length; out[n].weight = weight; }
      n++;
    }
  }
  return n;
}

static int sdp_fragments(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize, int sdpPrefixLength,
                         int maxMatches, uint32_t *frags, uint32_t capFrags) {
  const int small = wordSize < 5 ? wordSize : 5;                                   /* SDP_DETAILED_WORD_SIZE */
  const uint32_t P = (uint32_t)sdpPrefixLength;
  const uint32_t prefixLength = tLen < P ? tLen : P, suffixLength = (tLen - prefixLength) < P ? (tLen - prefixLength) : P;
  const uint32_t suffixPos = tLen - suffixLength;                                  /* prefix + middle lengths */
  const uint32_t qPrefixLength = qLen < P ? qLen : P, qSuffixLength = (qLen - qPrefixLength) < P ? (qLen - qPrefixLength) : P;
  const uint32_t qSuffixPos = qLen - qSuffixLength;
  TPos *lp = (TPos *)malloc(sizeof(TPos) * (prefixLength + 1)), *ls = (TPos *)malloc(sizeof(TPos) * (suffixLength + 1));
  TPos *lm = (TPos *)malloc(sizeof(TPos) * ((size_t)tLen + 1));
  const uint32_t np = tuple_list(t, prefixLength, small, lp), ns = tuple_list(t + suffixPos, suffixLength, small, ls);
  const uint32_t nm = tuple_list(t, tLen, wordSize, lm);
  uint32_t cap = capFrags, n = 0, i, m;
  Frag4 *all = (Frag4 *)malloc(sizeof(Frag4) * ((size_t)cap + 1));
  n = match_positions(q, qPrefixLength, small, lp, np, 0, 0, (uint32_t)small, (uint32_t)wordSize, all, n, cap, maxMatches);
  n = match_positions(q, qLen, wordSize, lm, nm, 0, 0, (uint32_t)wordSize, (uint32_t)wordSize, all, n, cap, maxMatches);
  n = match_positions(q + qSuffixPos, qSuffixLength, small, ls, ns, qSuffixPos, suffixPos, (uint32_t)small, (uint32_t)wordSize, all, n, cap, maxMatches);
  free(lp); free(ls); free(lm);
  if (n > cap) { free(all); return -1; }
  if (std_sort_xy(all, all + n)) { free(all); return ORC_SDP_UNPINNED; }
  m = 0;
  for (i = 0; i < n;) {                                                            /* keep the first of equal (x, y) */
    uint32_t j = i;
    all[m] = all[i];
    while (j < n && all[j].x == all[m].x && all[j].y == all[m].y) j++;
    m++; i = j;
  }
  for (i = 0; i < m; i++) { frags[4 * i] = all[i].x; frags[4 * i + 1] = all[i].y; frags[4 * i + 2] = all[i].length; frags[4 * i + 3] = all[i].weight; }
  free(all);
  return (int)m;
}

int orc_sdp_fragments(const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize, int sdpPrefixLength,
                      uint32_t *frags, uint32_t capFrags) {
  return sdp_fragments(q, qLen, t, tLen, wordSize, sdpPrefixLength, 0, frags, capFrags);
}

/* ------------------------------------------------------------------------------------------------------------------
 * SDPAlign as a whole (SDPAlign.h:95-637): fragment set, chain, chain -> blocks (:308-407), and the detailed part --
 * front extension (:409-474), SWAlign / recursive SDPAlign between chained blocks (:481-535), the last block and the tail
 * (:536-601), the Local shift (:604-612).  Blocks come back relative to (*qPos, *tPos) as the reference leaves them. */
typedef struct { uint32_t q, t, len; } Blk;
typedef struct { Blk *b; uint32_t n, cap; } BlkVec;
static void bv_push(BlkVec *v, Blk x) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 64; v->b = (Blk *)realloc(v->b, sizeof(Blk) * v->cap); }
  v->b[v->n++] = x;
}
/* SWAlign(qFragment, tFragment, ..., scoreFn, Global) through the restatement in orc_align.c; blocks appended raw */
static void sw_global(const orc_scorefn *fn, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, BlkVec *out,
                      uint32_t *qPos, uint32_t *tPos) {
  orc_job j; orc_result r;
  const uint32_t cap = qLen + tLen + 8;
  uint32_t *bl = (uint32_t *)malloc(sizeof(uint32_t) * 3 * cap), *gc = (uint32_t *)malloc(sizeof(uint32_t) * (cap + 1)), i;
  int32_t *gp = (int32_t *)malloc(sizeof(int32_t) * 4 * cap);
  memset(&j, 0, sizeof j);
  j.algo = ORC_SW; j.alignType = ORC_GLOBAL; j.q = q; j.qLen = qLen; j.t = t; j.tLen = tLen;
  orc_align(fn, &j, &r, bl, cap, gc, cap + 1, gp, 2 * cap);
  for (i = 0; i < r.nBlocks; i++) { Blk x; x.q = bl[3 * i]; x.t = bl[3 * i + 1]; x.len = bl[3 * i + 2]; bv_push(out, x); }
  *qPos = r.qPos; *tPos = r.tPos;
  free(bl); free(gc); free(gp);
}

static int sdp_align(const orc_scorefn *fn, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize,
                     int sdpIns, int sdpDel, float indelRate, int alignType, int detailed, int extendFront, int sdpPrefixLength,
                     int recurse, int noRecurseUnder, int maxMatches, BlkVec *out, uint32_t *qPos, uint32_t *tPos) {
  *qPos = 0; *tPos = 0;
  const uint32_t capF = 64u * (qLen + tLen) + 4096u;
  uint32_t *fr = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)capF);
  const int nF = sdp_fragments(q, qLen, t, tLen, wordSize, sdpPrefixLength, maxMatches, fr, capF);
  if (nF < 0) { free(fr); return nF; }
  if (nF == 0) { free(fr); return 0; }                       /* :264-269: needs at least one seed */
  int32_t *chain = (int32_t *)malloc(sizeof(int32_t) * ((size_t)nF + 1));
  const int nC = orc_sdp_chain(fr, (uint32_t)nF, qLen, (uint32_t)wordSize, sdpIns, sdpDel, fn->M[0], alignType, chain, (uint32_t)nF + 1);
  BlkVec ch = {0, 0, 0};
  int f;
  uint32_t b;
  /* :308-335 condense runs of fragments that advance by one in both sequences */
  for (f = 0; f < nC; f++) {
    const int startF = f;
    while (f < nC - 1 && fr[4 * chain[f]] == fr[4 * chain[f + 1]] - 1 && fr[4 * chain[f] + 1] == fr[4 * chain[f + 1] + 1] - 1) f++;
    Blk x; x.q = fr[4 * chain[startF]]; x.t = fr[4 * chain[startF] + 1];
    x.len = fr[4 * chain[f]] + fr[4 * chain[f] + 2] - fr[4 * chain[startF]];
    bv_push(&ch, x);
  }
  free(fr); free(chain);
  /* :349-358 a block may not run into the next one */
  for (b = 0; b + 1 < ch.n; b++) {
    if (ch.b[b].q + ch.b[b].len > ch.b[b + 1].q) ch.b[b].len = ch.b[b + 1].q - ch.b[b].q;
    if (ch.b[b].t + ch.b[b].len > ch.b[b + 1].t) ch.b[b].len = ch.b[b + 1].t - ch.b[b].t;
  }
  /* :373-407 drop empty blocks and blocks that sit off the diagonal of both neighbours */
  {
    uint8_t *good = (uint8_t *)malloc(ch.n + 1);
    uint32_t m = 0;
    for (b = 0; b < ch.n; b++) good[b] = ch.b[b].len != 0;
    for (b = 1; b + 1 < ch.n; b++) {
      const int prevDiag = abs(((int)ch.b[b].t - (int)ch.b[b].q) - ((int)ch.b[b - 1].t - (int)ch.b[b - 1].q));
      const uint32_t pdt = ch.b[b].t - ch.b[b - 1].t, pdq = ch.b[b].q - ch.b[b - 1].q;
      const int prevDist = (int)(pdt < pdq ? pdt : pdq);
      const int nextDiag = abs(((int)ch.b[b + 1].t - (int)ch.b[b + 1].q) - ((int)ch.b[b].t - (int)ch.b[b].q));
      const uint32_t ndt = ch.b[b + 1].t - ch.b[b].t, ndq = ch.b[b + 1].q - ch.b[b].q;
      const int nextDist = (int)(ndt < ndq ? ndt : ndq);
      if (prevDist * indelRate < prevDiag && nextDist * indelRate < nextDiag) good[b] = 0;
    }
    for (b = 0; b < ch.n; b++) if (good[b]) ch.b[m++] = ch.b[b];
    ch.n = m;
    free(good);
  }
  if (ch.n > 0) {
    const int sub = wordSize - 4 > 5 ? wordSize - 4 : 5;     /* max(wordSize-4, 5) */
    /* :412-474 front extension */
    if (ch.b[0].q > 0 && ch.b[0].t > 0 && (alignType == ORC_GLOBAL || extendFront)) {
      BlkVec fa = {0, 0, 0};
      uint32_t fq = 0, ft = 0, i;
      if (recurse == 0 && (uint32_t)(ch.b[0].q * ch.b[0].t) < (uint32_t)noRecurseUnder) sw_global(fn, q, ch.b[0].q, t, ch.b[0].t, &fa, &fq, &ft);
      else if (recurse != 0)
        /* as the reference (:456): smithWatermanAlignType (EndAnchored = 6) lands in the maxMatchesPerPosition slot */
        sdp_align(fn, q, ch.b[0].q, t, ch.b[0].t, sub, sdpIns, sdpDel, indelRate, ORC_GLOBAL, detailed, extendFront, sdpPrefixLength,
                  recurse - 1, noRecurseUnder, ORC_ENDANCHORED, &fa, &fq, &ft);
      for (i = 0; i < fa.n; i++) { Blk x = fa.b[i]; x.t += ft; x.q += fq; bv_push(out, x); }
      free(fa.b);
    }
    /* :481-535 the chained blocks and what lies between them */
    for (b = 0; b + 1 < ch.n; b++) {
      bv_push(out, ch.b[b]);
      const uint32_t qo = ch.b[b].q + ch.b[b].len, to = ch.b[b].t + ch.b[b].len;
      const uint32_t ql = ch.b[b + 1].q - qo, tl = ch.b[b + 1].t - to;
      if (ql > 0 && tl > 0 && detailed) {
        BlkVec fa = {0, 0, 0};
        uint32_t fq = 0, ft = 0, i;
        if ((uint32_t)(ql * tl) < (uint32_t)noRecurseUnder) sw_global(fn, q + qo, ql, t + to, tl, &fa, &fq, &ft);
        else if (recurse != 0)
          sdp_align(fn, q + qo, ql, t + to, tl, sub, sdpIns, sdpDel, indelRate, ORC_GLOBAL, detailed, 0, 0, recurse - 1, noRecurseUnder, 0,
                    &fa, &fq, &ft);
        /* :523-524 the fragment's own qPos / tPos are reset, its blocks are used as they are */
        for (i = 0; i < fa.n; i++) { Blk x = fa.b[i]; x.q += qo; x.t += to; bv_push(out, x); }
        free(fa.b);
      }
    }
    /* :536-601 the last block, and the tail when front extension is on */
    if (alignType == ORC_GLOBAL || alignType == ORC_LOCAL) {
      const Blk last = ch.b[ch.n - 1];
      bv_push(out, last);
      if (alignType == ORC_GLOBAL || extendFront) {
        const uint32_t qo = last.q + last.len, to = last.t + last.len;
        const uint32_t ql = qLen - qo, tl = tLen - to;
        if (ql > 0 && tl > 0 && extendFront) {
          BlkVec fa = {0, 0, 0};
          uint32_t fq = 0, ft = 0, i;
          const int half = wordSize / 2 > 5 ? wordSize / 2 : 5;
          if (recurse == 0 && (uint32_t)(ql * tl) < (uint32_t)noRecurseUnder) sw_global(fn, q + qo, ql, t + to, tl, &fa, &fq, &ft);
          else if (recurse != 0)
            sdp_align(fn, q + qo, ql, t + to, tl, half, sdpIns, sdpDel, indelRate, ORC_GLOBAL, detailed, extendFront, sdpPrefixLength,
                      recurse - 1, noRecurseUnder, maxMatches, &fa, &fq, &ft);
          for (i = 0; i < fa.n; i++) { Blk x = fa.b[i]; x.q += qo; x.t += to; bv_push(out, x); }
          free(fa.b);
        }
      }
    }
  }
  free(ch.b);
  return 0;
}

int orc_sdp_align(const orc_scorefn *fn, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize,
                  int sdpIns, int sdpDel, float indelRate, int alignType, int detailed, int extendFront, int sdpPrefixLength,
                  int recurse, int noRecurseUnder, uint32_t *blocks, uint32_t capBlocks, uint32_t *qPos, uint32_t *tPos) {
  BlkVec out = {0, 0, 0};
  uint32_t i;
  const int rc = sdp_align(fn, q, qLen, t, tLen, wordSize, sdpIns, sdpDel, indelRate, alignType, detailed, extendFront, sdpPrefixLength,
                           recurse, noRecurseUnder, 0, &out, qPos, tPos);
  if (rc < 0) { free(out.b); return rc; }
  if (alignType == ORC_LOCAL && out.n > 0) {                 /* :604-612 */
    *tPos = out.b[0].t; *qPos = out.b[0].q;
    for (i = 0; i < out.n; i++) { out.b[i].q -= *qPos; out.b[i].t -= *tPos; }
  }
  if (out.n > capBlocks) { free(out.b); return -1; }
  for (i = 0; i < out.n; i++) { blocks[3 * i] = out.b[i].q; blocks[3 * i + 1] = out.b[i].t; blocks[3 * i + 2] = out.b[i].len; }
  i = out.n;
  free(out.b);
  return (int)i;
}
