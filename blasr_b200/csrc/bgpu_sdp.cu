// bgpu_sdp.cu -- SDPAlign on the device: the step that produces the guide the refinement consumes.
//
// Reference semantics restated (not translated):
//   SDPAlign                       common/algorithms/alignment/SDPAlign.h:95-637
//     fragment set                 :133-262   (SequenceToTupleList tuples/DNATuple.h:309-358, StoreMatchingPositions
//                                              tuples/TupleMatching.h:37-62, FindAll tuples/TupleList.h:88-101)
//     chain                        SDPLongestCommonSubsequence sdp/SparseDynamicProgramming.h:71-322, StoreAbove :51-69,
//                                  IndelPenalty :27-49, SDPSet sdp/SDPSet.h:16-120, Fragment order sdp/SDPFragment.h:62-93
//     chain -> blocks              :308-407
//     front extension / gap fills / tail   :409-601 (SWAlign.h:18-389 Global for boxes under noRecurseUnder cells,
//                                  SDPAlign itself with a smaller word otherwise), Local shift :604-612
//
// B200 mapping, first device version: ONE THREAD PER JOB (lane 0 of a warp).  Everything after the k-mer matching is a sequential algorithm whose
// result depends on its exact order of operations -- the (x, y) sort is libstdc++'s introsort (median-of-three quicksort,
// heapsort escape, final insertion sort) because the survivor among equal (x, y) fragments of different length is whichever
// that unstable sort leaves first; the chain is a sweep over two ordered sets -- so a job is walked by one thread and the
// chip is filled with jobs (persistent threads pull jobs from a counter; each owns a slice of a scratch arena and allocates
// from it stack-wise, recursion included).  k-mer matching needs no sorted tuple list: the target's k-mers go into an
// open-addressing table whose per-k-mer position lists are built from the last position to the first, so they come out
// ascending, which is the order FindAll enumerates.  A job that outgrows its arena slice comes back BGPU_JOB_RANGE.
#include "bgpu_common.cuh"

namespace bgpu {
namespace sdp {

struct F4 { uint32_t x, y, length, weight; };
struct Blk { uint32_t q, t, len; };
struct Args {                 // SDPAlign's parameter list as blasr fills it (Blasr.cpp:1716-1722, :1080-1090)
  int M[25]; int ins, del;    // DistanceMatrixScoreFunction: SWAlign gap fills; M[0] is the chain's `match`
  int sdpIns, sdpDel; float indelRate;
};
struct Arena {
  uint8_t *base; size_t cap, top; bool oom;
  __device__ void *alloc(size_t bytes) {
    const size_t at = (top + 15) & ~(size_t)15;
    if (at + bytes > cap) { oom = true; return nullptr; }
    top = at + bytes;
    return base + at;
  }
};

__device__ __forceinline__ int base2(uint8_t c) {   // TwoBit where ThreeBit <= 3 (NucConversion.h:7-84), else -1
  const uint8_t b = base_code(c);
  return b <= 3 ? (int)b : -1;
}

// ---- std::sort on (x, y), libstdc++ (bits/stl_algo.h, bits/stl_heap.h): which of two equal keys ends up first is part of
//      the reference's result (SDPAlign.h:249-262 keeps the first)
__device__ __forceinline__ bool less_xy(const F4 &a, const F4 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }
__device__ __forceinline__ void swap4(F4 &a, F4 &b) { const F4 t = a; a = b; b = t; }
__device__ void move_median_to_first(F4 *result, F4 *a, F4 *b, F4 *c) {
  if (less_xy(*a, *b)) {
    if (less_xy(*b, *c)) swap4(*result, *b);
    else if (less_xy(*a, *c)) swap4(*result, *c);
    else swap4(*result, *a);
  } else if (less_xy(*a, *c)) swap4(*result, *a);
  else if (less_xy(*b, *c)) swap4(*result, *c);
  else swap4(*result, *b);
}
__device__ F4 *unguarded_partition(F4 *first, F4 *last, F4 *pivot) {
  for (;;) {
    while (less_xy(*first, *pivot)) ++first;
    --last;
    while (less_xy(*pivot, *last)) --last;
    if (!(first < last)) return first;
    swap4(*first, *last);
    ++first;
  }
}
__device__ void push_heap_(F4 *first, long hole, long top, F4 value) {
  long parent = (hole - 1) / 2;
  while (hole > top && less_xy(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
  first[hole] = value;
}
__device__ void adjust_heap(F4 *first, long hole, long len, F4 value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less_xy(first[child], first[child - 1])) child--;
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
__device__ void heap_sort(F4 *first, F4 *last) {
  const long len = last - first;
  if (len >= 2) for (long parent = (len - 2) / 2;; parent--) { adjust_heap(first, parent, len, first[parent]); if (parent == 0) break; }
  while (last - first > 1) { --last; const F4 value = *last; *last = *first; adjust_heap(first, 0, last - first, value); }
}
// the recursion of __introsort_loop on the right part, as an explicit stack of (cut, last, depth): at most 2 log2 n deep
__device__ void introsort_loop(F4 *first, F4 *last, long depth) {
  struct Fr { F4 *first, *last; long depth; } st[72];
  int sp = 0;
  st[sp++] = Fr{first, last, depth};
  while (sp) {
    Fr f = st[--sp];
    // the reference recurses into [cut, last) FIRST and then continues with [first, cut): the right parts are finished
    // in order before the left part is touched, so the left part is pushed below the right one
    while (f.last - f.first > 16) {
      if (f.depth == 0) { heap_sort(f.first, f.last); break; }
      --f.depth;
      F4 *mid = f.first + (f.last - f.first) / 2;
      move_median_to_first(f.first, f.first + 1, mid, f.last - 1);
      F4 *cut = unguarded_partition(f.first + 1, f.last, f.first);
      // continue with the RIGHT part now, come back to the left part later: partitions of disjoint ranges commute, the
      // result is the reference's
      if (sp < 72) st[sp++] = Fr{f.first, cut, f.depth};
      f.first = cut;
    }
  }
}
__device__ void unguarded_linear_insert(F4 *last) {
  const F4 val = *last;
  F4 *next = last - 1;
  while (less_xy(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
__device__ void insertion_sort(F4 *first, F4 *last) {
  if (first == last) return;
  for (F4 *i = first + 1; i != last; ++i) {
    if (less_xy(*i, *first)) { const F4 val = *i; for (F4 *p = i; p != first; --p) *p = *(p - 1); *first = val; }
    else unguarded_linear_insert(i);
  }
}
__device__ void std_sort_xy(F4 *first, F4 *last) {
  const long n = last - first;
  long lg = 0;
  if (first == last) return;
  while ((1L << (lg + 1)) <= n) lg++;                      // std::__lg
  introsort_loop(first, last, lg * 2);
  if (n > 16) { insertion_sort(first, first + 16); for (F4 *i = first + 16; i != last; ++i) unguarded_linear_insert(i); }
  else insertion_sort(first, last);
}

// ---- k-mer matches of q[0, qLen) against t[0, tLen): (s + xOff, pos + yOff) for every target position holding the query's
//      k-mer at s, s ascending, positions ascending (StoreMatchingPositions over a (tuple, pos)-sorted list).  Appends to
//      out[n ...), returns the new count (which may exceed cap: the caller turns that into an overflow).
__device__ uint32_t match_set(Arena &A, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int k, uint32_t xOff,
                              uint32_t yOff, uint32_t length, uint32_t weight, F4 *out, uint32_t n, uint32_t cap, int maxMatches) {
  if (k <= 0 || tLen < (uint32_t)k || qLen < (uint32_t)k) return n;
  const size_t mark = A.top;
  uint32_t H = 16;
  while (H < 2u * tLen) H <<= 1;
  uint32_t *keys = (uint32_t *)A.alloc(sizeof(uint32_t) * H);
  int32_t *head = (int32_t *)A.alloc(sizeof(int32_t) * H), *next = (int32_t *)A.alloc(sizeof(int32_t) * tLen);
  if (A.oom) { A.top = mark; return n; }
  for (uint32_t i = 0; i < H; i++) keys[i] = 0;
  const uint32_t mask = k >= 16 ? 0xffffffffu : ((1u << (2 * k)) - 1u);
  int shiftH = 0; while ((1u << shiftH) < H) shiftH++;
  auto slot_of = [&](uint32_t v) { return (uint32_t)((v * 2654435761u) >> (32 - shiftH)); };
  {   // target k-mers, last position first: every list comes out ascending
    uint32_t v = 0, run = 0;
    for (uint32_t pp = tLen; pp-- > 0;) {
      const int b = base2(t[pp]);
      if (b < 0) { run = 0; v = 0; continue; }
      v = ((v << 2) | (uint32_t)b) & mask; run++;
      if (run >= (uint32_t)k) {
        uint32_t s = slot_of(v);
        while (keys[s] != 0 && keys[s] != v + 1) s = (s + 1) & (H - 1);
        if (keys[s] == 0) { keys[s] = v + 1; next[pp] = -1; } else next[pp] = head[s];
        head[s] = (int32_t)pp;
      }
    }
  }
  {   // the query's k-mers, first position first (the leftmost base sits in the lowest two bits, DNATuple.h:55-83)
    uint32_t v = 0, run = 0;
    for (uint32_t e = 0; e < qLen; e++) {
      const int b = base2(q[e]);
      if (b < 0) { run = 0; v = 0; continue; }
      v = (v >> 2) | ((uint32_t)b << (2 * (k - 1))); run++;
      if (run < (uint32_t)k) continue;
      const uint32_t s0 = e + 1 - (uint32_t)k;
      uint32_t s = slot_of(v);
      while (keys[s] != 0 && keys[s] != v + 1) s = (s + 1) & (H - 1);
      if (keys[s] == 0) continue;
      if (maxMatches != 0) {                                 // positions with more matches than that are skipped (:50)
        long cnt = 0;
        for (int32_t p = head[s]; p >= 0; p = next[p]) cnt++;
        if (cnt > (long)maxMatches) continue;
      }
      for (int32_t p = head[s]; p >= 0; p = next[p]) {
        if (n < cap) out[n] = F4{s0 + xOff, (uint32_t)p + yOff, length, weight};
        n++;
      }
    }
  }
  A.top = mark;
  return n;
}

// the fragment set of SDPAlign.h:133-262; returns the count after de-duplication, or -1 (arena too small)
__device__ int fragments(Arena &A, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, int wordSize, int sdpPrefixLength,
                         int maxMatches, F4 *all, uint32_t cap) {
  const int small = wordSize < 5 ? wordSize : 5;                                   // SDP_DETAILED_WORD_SIZE
  const uint32_t P = (uint32_t)sdpPrefixLength;
  const uint32_t prefixLength = tLen < P ? tLen : P, suffixLength = (tLen - prefixLength) < P ? (tLen - prefixLength) : P;
  const uint32_t suffixPos = tLen - suffixLength;
  const uint32_t qPrefixLength = qLen < P ? qLen : P, qSuffixLength = (qLen - qPrefixLength) < P ? (qLen - qPrefixLength) : P;
  const uint32_t qSuffixPos = qLen - qSuffixLength;
  uint32_t n = 0;
  n = match_set(A, q, qPrefixLength, t, prefixLength, small, 0, 0, (uint32_t)small, (uint32_t)wordSize, all, n, cap, maxMatches);
  n = match_set(A, q, qLen, t, tLen, wordSize, 0, 0, (uint32_t)wordSize, (uint32_t)wordSize, all, n, cap, maxMatches);
  n = match_set(A, q + qSuffixPos, qSuffixLength, t + suffixPos, suffixLength, small, qSuffixPos, suffixPos, (uint32_t)small,
                (uint32_t)wordSize, all, n, cap, maxMatches);
  if (A.oom || n > cap) return -1;
  std_sort_xy(all, all + n);
  uint32_t m = 0;
  for (uint32_t i = 0; i < n;) {                                                   // keep the first of equal (x, y)
    uint32_t j = i;
    all[m] = all[i];
    while (j < n && all[j].x == all[m].x && all[j].y == all[m].y) j++;
    m++; i = j;
  }
  return (int)m;
}

// ---- the chain (SparseDynamicProgramming.h:71-322) over a fragment set sorted by (x, y) with unique keys
struct Col { int col, opt; };
struct Swept { uint32_t x, y; int cost, index; };
__device__ __forceinline__ bool swept_less(uint32_t ax, uint32_t ay, uint32_t bx, uint32_t by) {   // SDPFragment.h:78-93
  const int da = (int)(ay - ax), db = (int)by - (int)bx;
  if (da < db) return true;
  if (da == db) return ax < bx;
  return false;
}
__device__ __forceinline__ int indel_penalty(int x1, int y1, int x2, int y2, int insertion, int deletion) {   // :27-49
  const int drift = (x1 - y1) - (x2 - y2);
  if (drift > 0) return (int)((1.0 * drift) * insertion);
  if (drift < 0) return (int)((-1.0 * drift) * deletion);
  return 0;
}
__device__ int swept_lower_bound(const Swept *s, int n, uint32_t x, uint32_t y) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (swept_less(s[mid].x, s[mid].y, x, y)) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ int col_lower_bound(const Col *c, int n, int col) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid].col < col) lo = mid + 1; else hi = mid; }
  return lo;
}
// returns the chain length (indices into f, first fragment first) or -1 (arena too small)
__device__ int chain_of(Arena &A, const F4 *f, uint32_t n, uint32_t queryLength, uint32_t fragmentLength, int insertion, int deletion,
                        int match, int alignType, int32_t *chain) {
  if (n < 1) return 0;
  const size_t mark = A.top;
  int *cost = (int *)A.alloc(sizeof(int) * n), *chainPrev = (int *)A.alloc(sizeof(int) * n), *above = (int *)A.alloc(sizeof(int) * n);
  uint32_t *chainLength = (uint32_t *)A.alloc(sizeof(uint32_t) * n);
  Col *cols = (Col *)A.alloc(sizeof(Col) * (n + 1));
  Swept *sw = (Swept *)A.alloc(sizeof(Swept) * (n + 1));
  uint32_t *ord = (uint32_t *)A.alloc(sizeof(uint32_t) * n);
  if (A.oom) { A.top = mark; return -1; }
  for (uint32_t i = 0; i < n; i++) { cost[i] = 0; chainPrev[i] = 0; above[i] = -1; chainLength[i] = 0; ord[i] = i; }
  {   // StoreAbove :51-69: neighbours in (y, x) order (a heapsort of the indices: the key is unique)
    auto lessYX = [&](uint32_t a, uint32_t b) { return f[a].y < f[b].y || (f[a].y == f[b].y && f[a].x < f[b].x); };
    auto sift = [&](uint32_t start, uint32_t end) {
      uint32_t root = start;
      for (;;) {
        uint32_t child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && lessYX(ord[child], ord[child + 1])) child++;
        if (!lessYX(ord[root], ord[child])) break;
        const uint32_t tmp = ord[root]; ord[root] = ord[child]; ord[child] = tmp;
        root = child;
      }
    };
    for (uint32_t s = n / 2; s-- > 0;) sift(s, n);
    for (uint32_t e = n; e-- > 1;) { const uint32_t tmp = ord[0]; ord[0] = ord[e]; ord[e] = tmp; sift(0, e); }
    for (uint32_t i = 1; i < n; i++) {
      const F4 &p = f[ord[i - 1]], &c = f[ord[i]];
      if (p.x <= c.x && p.x + p.length > c.x && p.y < c.y) above[ord[i]] = (int)ord[i - 1];
    }
  }
  int nCols = 0, nSw = 0;
  uint32_t sweepRow = f[0].x, fSweep = 0, fTrail = 0, maxChainLength = 0;
  int maxChainFragment = -1, minFragmentCost = INT_MAX, minFragmentIndex = -1;
  for (; sweepRow < queryLength + fragmentLength; sweepRow++) {                    // :108
    const uint32_t startF = fSweep;
    while (fSweep < n && f[fSweep].x == sweepRow) {
      const F4 &c = f[fSweep];
      int cp = INT_MAX, cl = INT_MAX, ca = INT_MAX, predOpt = -1, predIndex = -1;
      bool foundPrev = false;
      if (nCols > 0) {                                      // colSet.Predecessor: the greatest col <= y (SDPSet.h:95-120)
        int it = col_lower_bound(cols, nCols, (int)c.y);
        bool have = false;
        if (it < nCols && cols[it].col == (int)c.y) have = true;
        else { if (it != 0) --it; if (!((int)c.y < cols[it].col)) have = true; }
        if (have) {
          predOpt = cols[it].opt;
          const int dist = abs((int)(c.x + c.y) - (int)(f[predOpt].x + f[predOpt].y));
          cp = (int)((uint32_t)cost[predOpt] + (uint32_t)(int)sqrt((double)dist) - c.length);   // :133-136
          foundPrev = true;
        }
      }
      if (nSw > 0) {                                        // sweepSet.Predecessor: the greatest (diagonal, x) <= this one's
        int it = swept_lower_bound(sw, nSw, c.x, c.y);
        bool have = false;
        if (it < nSw && !swept_less(c.x, c.y, sw[it].x, sw[it].y)) have = true;
        else { if (it != 0) --it; if (!swept_less(c.x, c.y, sw[it].x, sw[it].y)) have = true; }
        if (have) {
          const Swept &p = sw[it];
          const int overlap = (int)(fragmentLength - (c.y - p.y)) * match;                     // :157
          cl = p.cost + (overlap < 0 ? overlap : 0) + indel_penalty((int)c.x, (int)c.y, (int)p.x, (int)p.y, insertion, deletion);
          predIndex = p.index;
          foundPrev = true;
        }
      }
      if (above[fSweep] >= 0) {                                                                // :164-175
        const int a = above[fSweep];
        ca = (int)((uint32_t)cost[a] + (fragmentLength - (uint32_t)(int)(c.y - f[a].y)) * (uint32_t)match +
                   (uint32_t)indel_penalty((int)c.x, (int)c.y, (int)f[a].x, (int)f[a].y, insertion, deletion));
        foundPrev = true;
      }
      int minCost = cl < ca ? cl : ca;
      minCost = cp < minCost ? cp : minCost;
      if (foundPrev && (alignType == BGPU_GLOBAL || (alignType == BGPU_LOCAL && minCost < 0))) {
        cost[fSweep] = (int)((uint32_t)minCost - c.weight);
        if (minCost == cp) chainPrev[fSweep] = predOpt;
        else if (minCost == cl) chainPrev[fSweep] = predIndex;
        else if (minCost == ca) chainPrev[fSweep] = above[fSweep];
        chainLength[fSweep] = chainLength[chainPrev[fSweep]] + 1;
      } else if (alignType == BGPU_GLOBAL) {
        chainPrev[fSweep] = -1;
        cost[fSweep] = (int)((c.x + c.y) * (uint32_t)deletion + fragmentLength * (uint32_t)match - c.weight);   // :211
        chainLength[fSweep] = 1;
      } else if (alignType == BGPU_LOCAL) {
        chainPrev[fSweep] = -1;
        cost[fSweep] = (int)(fragmentLength * (uint32_t)match - c.weight);
        chainLength[fSweep] = 1;
      }
      if (minFragmentCost > cost[fSweep]) { minFragmentCost = cost[fSweep]; minFragmentIndex = (int)fSweep; }
      if (chainLength[fSweep] > maxChainLength) { maxChainLength = chainLength[fSweep]; maxChainFragment = (int)fSweep; }
      fSweep++;
    }
    for (fSweep = startF; fSweep < n && f[fSweep].x == sweepRow; fSweep++) {        // the row's fragments enter the sweep set
      int it = swept_lower_bound(sw, nSw, f[fSweep].x, f[fSweep].y);
      if (!(it < nSw && !swept_less(f[fSweep].x, f[fSweep].y, sw[it].x, sw[it].y))) {
        for (int m = nSw; m > it; m--) sw[m] = sw[m - 1];
        nSw++;
      }
      sw[it] = Swept{f[fSweep].x, f[fSweep].y, cost[fSweep], (int)fSweep};
    }
    if (sweepRow >= fragmentLength + 1) {       // fragments fragmentLength + 1 rows back leave the sweep set (:240-300)
      const uint32_t trailRow = sweepRow - fragmentLength - 1;
      while (fTrail < n && f[fTrail].x == trailRow) {
        const int y = (int)f[fTrail].y;
        int it = col_lower_bound(cols, nCols, y);
        bool storeCol;
        if (it < nCols && cols[it].col == y) storeCol = cost[cols[it].opt] < cost[fTrail];   // as the reference (:258-262): the
                                                        // existing entry is replaced when it is the CHEAPER one
        else storeCol = true;
        if (storeCol) {
          if (it < nCols && cols[it].col == y) cols[it].opt = (int)fTrail;
          else { for (int m = nCols; m > it; m--) cols[m] = cols[m - 1]; cols[it] = Col{y, (int)fTrail}; nCols++; }
          // Successor answers "none" for sets of fewer than two elements (SDPSet.h:79-81)
          while (nCols >= 2 && it + 1 < nCols && cost[cols[it + 1].opt] > cost[fTrail]) {
            for (int m = it + 1; m + 1 < nCols; m++) cols[m] = cols[m + 1];
            nCols--;
          }
        }
        {   // sweepSet.Delete
          const int s = swept_lower_bound(sw, nSw, f[fTrail].x, f[fTrail].y);
          if (s < nSw && sw[s].x == f[fTrail].x && sw[s].y == f[fTrail].y) { for (int m = s; m + 1 < nSw; m++) sw[m] = sw[m + 1]; nSw--; }
        }
        ++fTrail;
      }
    }
  }
  if (alignType == BGPU_LOCAL) maxChainFragment = minFragmentIndex;
  int len = 0;
  for (int k = maxChainFragment; k != -1; k = chainPrev[k]) len++;
  int i = len;
  for (int k = maxChainFragment; k != -1; k = chainPrev[k]) chain[--i] = k;
  A.top = mark;
  return len;
}

// ---- SWAlign(..., Global) (SWAlign.h:18-389) of a gap-fill box: blocks appended to out (positions inside the box)
__device__ uint32_t sw_global(Arena &A, const Args &a, const uint8_t *q, uint32_t qLen, const uint8_t *t, uint32_t tLen, Blk *out,
                              uint32_t nOut, uint32_t capOut) {
  const size_t mark = A.top;
  const int64_t nCols = (int64_t)tLen + 1, nRows = (int64_t)qLen + 1;
  int *S = (int *)A.alloc(sizeof(int) * (size_t)(nRows * nCols));
  uint8_t *P = (uint8_t *)A.alloc((size_t)(nRows * nCols));
  if (A.oom) { A.top = mark; return nOut; }
  enum { DIAG = 0, UP = 1, LEFT = 2 };
  for (int64_t c = 0; c < nCols; c++) { S[c] = (int)(a.del * c); P[c] = LEFT; }                 // :49-138, Global
  for (int64_t r = 0; r < nRows; r++) { S[r * nCols] = (int)(a.ins * r); P[r * nCols] = UP; }
  P[0] = DIAG;                                                                                 // :140
  for (uint32_t r = 0; r < qLen; r++) {
    const int qc = base_code(q[r]);
    for (uint32_t c = 0; c < tLen; c++) {
      const int ms = a.M[qc * 5 + base_code(t[c])] + S[r * nCols + c];
      const int qg = S[r * nCols + c + 1] + a.ins;                                             // :166
      const int tg = S[(r + 1) * nCols + c] + a.del;                                           // :167
      const int best = min(ms, min(qg, tg));
      S[(r + 1) * nCols + c + 1] = best;
      P[(r + 1) * nCols + c + 1] = best == ms ? DIAG : (best == qg ? UP : LEFT);               // Diagonal > Up > Left :196-207
    }
  }
  // traceback from (qLen, tLen) to the origin (:324-353); the path is walked backwards, so the blocks are produced last first
  const uint32_t first = nOut;
  int64_t r = qLen, c = tLen;
  while (r > 0 || c > 0) {
    const uint8_t ar = P[r * nCols + c];
    if (ar == DIAG) {
      uint32_t len = 0;
      while ((r > 0 || c > 0) && P[r * nCols + c] == DIAG) { len++; r--; c--; }
      if (nOut < capOut) out[nOut] = Blk{(uint32_t)r, (uint32_t)c, len};
      nOut++;
    } else if (ar == UP) r--;
    else c--;
  }
  for (uint32_t i = first, j = min(nOut, capOut); i + 1 < j; i++, j--) { const Blk x = out[i]; out[i] = out[j - 1]; out[j - 1] = x; }
  A.top = mark;
  return nOut;
}

// ---- SDPAlign (SDPAlign.h:95-637); blocks appended to out relative to the sequences' starts; returns the new count, sets
//      A.oom / ovf when the arena slice or out is too small
struct Frame {
  const uint8_t *q, *t; uint32_t qLen, tLen; int wordSize, alignType, detailed, extendFront, sdpPrefixLength, recurse, noRecurseUnder, maxMatches;
};
__device__ uint32_t sdp_align(Arena &A, const Args &a, const Frame fr, Blk *out, uint32_t nOut, const uint32_t capOut, bool &ovf) {
  const size_t mark = A.top;
  const uint32_t qLen = fr.qLen, tLen = fr.tLen;
  const uint8_t *q = fr.q, *t = fr.t;
  const int wordSize = fr.wordSize;
  if (wordSize > 15) { ovf = true; return nOut; }
  // fragment capacity: what is left of the slice, less what the chain will need (56 bytes per fragment) and a reserve
  const size_t left = A.cap > A.top ? A.cap - A.top : 0;
  const size_t hashNeed = 40 * (size_t)tLen + 4096;     // the k-mer table of the whole target is built above the fragment array
  if (left < hashNeed + 8192) { A.oom = true; return nOut; }
  uint32_t capF = (uint32_t)min((size_t)0x7fffffff / 80, (left - hashNeed) / 80);
  F4 *fs = (F4 *)A.alloc(sizeof(F4) * (size_t)capF);
  if (A.oom) return nOut;
  const int nF = fragments(A, q, qLen, t, tLen, wordSize, fr.sdpPrefixLength, fr.maxMatches, fs, capF);
  if (nF < 0) { A.oom = true; A.top = mark; return nOut; }
  if (nF == 0) { A.top = mark; return nOut; }                // :264-269: needs at least one seed
  A.top = (size_t)((uint8_t *)(fs + nF) - A.base);           // give back the unused tail of the fragment array
  int32_t *chain = (int32_t *)A.alloc(sizeof(int32_t) * ((size_t)nF + 1));
  if (A.oom) { A.top = mark; return nOut; }
  const int nC = chain_of(A, fs, (uint32_t)nF, qLen, (uint32_t)wordSize, a.sdpIns, a.sdpDel, a.M[0], fr.alignType, chain);
  if (nC < 0) { A.oom = true; A.top = mark; return nOut; }
  // :308-335 condense runs of fragments that advance by one in both sequences (written over the front of a new array)
  Blk *ch = (Blk *)A.alloc(sizeof(Blk) * ((size_t)nC + 1));
  if (A.oom) { A.top = mark; return nOut; }
  uint32_t nCh = 0;
  for (int f = 0; f < nC; f++) {
    const int startF = f;
    while (f < nC - 1 && fs[chain[f]].x == fs[chain[f + 1]].x - 1 && fs[chain[f]].y == fs[chain[f + 1]].y - 1) f++;
    ch[nCh++] = Blk{fs[chain[startF]].x, fs[chain[startF]].y, fs[chain[f]].x + fs[chain[f]].length - fs[chain[startF]].x};
  }
  // :349-358 a block may not run into the next one
  for (uint32_t b = 0; b + 1 < nCh; b++) {
    if (ch[b].q + ch[b].len > ch[b + 1].q) ch[b].len = ch[b + 1].q - ch[b].q;
    if (ch[b].t + ch[b].len > ch[b + 1].t) ch[b].len = ch[b + 1].t - ch[b].t;
  }
  {   // :373-407 drop empty blocks and blocks that sit off the diagonal of both neighbours (decided on the unfiltered list)
    uint8_t *good = (uint8_t *)A.alloc((size_t)nCh + 1);
    if (A.oom) { A.top = mark; return nOut; }
    for (uint32_t b = 0; b < nCh; b++) good[b] = ch[b].len != 0;
    for (uint32_t b = 1; b + 1 < nCh; b++) {
      const int prevDiag = abs(((int)ch[b].t - (int)ch[b].q) - ((int)ch[b - 1].t - (int)ch[b - 1].q));
      const uint32_t pdt = ch[b].t - ch[b - 1].t, pdq = ch[b].q - ch[b - 1].q;
      const int prevDist = (int)(pdt < pdq ? pdt : pdq);
      const int nextDiag = abs(((int)ch[b + 1].t - (int)ch[b + 1].q) - ((int)ch[b].t - (int)ch[b].q));
      const uint32_t ndt = ch[b + 1].t - ch[b].t, ndq = ch[b + 1].q - ch[b].q;
      const int nextDist = (int)(ndt < ndq ? ndt : ndq);
      if (prevDist * a.indelRate < prevDiag && nextDist * a.indelRate < nextDiag) good[b] = 0;
    }
    uint32_t m = 0;
    for (uint32_t b = 0; b < nCh; b++) if (good[b]) ch[m++] = ch[b];
    nCh = m;
  }
  // the chained blocks survive below; everything between them and the fragment array is dead now: move them down
  {
    Blk *dst = (Blk *)(A.base + ((mark + 15) & ~(size_t)15));
    for (uint32_t b = 0; b < nCh; b++) dst[b] = ch[b];
    ch = dst;
    A.top = (size_t)((uint8_t *)(ch + nCh) - A.base);
  }
  auto push = [&](const Blk &x) { if (nOut < capOut) out[nOut] = x; else ovf = true; nOut++; };
  // a sub-alignment's blocks land at out[nOut ...) relative to its own box: shift them by the box's origin
  auto sub = [&](const uint32_t qo, const uint32_t to, const uint32_t ql, const uint32_t tl, const int word, const int extendFront,
                 const int prefix, const int maxMatches, const bool swOnly, const bool swAllowed) {
    const uint32_t first = nOut;
    if (swAllowed && (uint32_t)(ql * tl) < (uint32_t)fr.noRecurseUnder) nOut = sw_global(A, a, q + qo, ql, t + to, tl, out, nOut, capOut);
    else if (!swOnly && fr.recurse != 0) {
      Frame s = fr;
      s.q = q + qo; s.t = t + to; s.qLen = ql; s.tLen = tl; s.wordSize = word; s.alignType = BGPU_GLOBAL; s.extendFront = extendFront;
      s.sdpPrefixLength = prefix; s.recurse = fr.recurse - 1; s.maxMatches = maxMatches;
      nOut = sdp_align(A, a, s, out, nOut, capOut, ovf);
    }
    if (nOut > capOut) ovf = true;
    for (uint32_t i = first; i < nOut && i < capOut; i++) { out[i].q += qo; out[i].t += to; }
  };
  if (nCh > 0) {
    const int subWord = wordSize - 4 > 5 ? wordSize - 4 : 5;     // max(wordSize - 4, 5)
    // :412-474 front extension: SWAlign only when recursion is exhausted, else SDPAlign with the reference's argument slip
    // (:456: smithWatermanAlignType = EndAnchored = 6 lands in the maxMatchesPerPosition slot)
    if (ch[0].q > 0 && ch[0].t > 0 && (fr.alignType == BGPU_GLOBAL || fr.extendFront)) {
      if (fr.recurse == 0) sub(0, 0, ch[0].q, ch[0].t, subWord, fr.extendFront, fr.sdpPrefixLength, BGPU_ENDANCHORED, true, true);
      else sub(0, 0, ch[0].q, ch[0].t, subWord, fr.extendFront, fr.sdpPrefixLength, BGPU_ENDANCHORED, false, false);
    }
    // :481-535 the chained blocks and what lies between them
    for (uint32_t b = 0; b + 1 < nCh; b++) {
      push(ch[b]);
      const uint32_t qo = ch[b].q + ch[b].len, to = ch[b].t + ch[b].len;
      const uint32_t ql = ch[b + 1].q - qo, tl = ch[b + 1].t - to;
      if (ql > 0 && tl > 0 && fr.detailed) sub(qo, to, ql, tl, subWord, 0, 0, 0, false, true);
    }
    // :536-601 the last block, and the tail when front extension is on
    if (fr.alignType == BGPU_GLOBAL || fr.alignType == BGPU_LOCAL) {
      const Blk last = ch[nCh - 1];
      push(last);
      if (fr.alignType == BGPU_GLOBAL || fr.extendFront) {
        const uint32_t qo = last.q + last.len, to = last.t + last.len;
        const uint32_t ql = qLen - qo, tl = tLen - to;
        if (ql > 0 && tl > 0 && fr.extendFront) {
          const int half = wordSize / 2 > 5 ? wordSize / 2 : 5;
          if (fr.recurse == 0) sub(qo, to, ql, tl, half, fr.extendFront, fr.sdpPrefixLength, fr.maxMatches, true, true);
          else sub(qo, to, ql, tl, half, fr.extendFront, fr.sdpPrefixLength, fr.maxMatches, false, false);
        }
      }
    }
  }
  A.top = mark;
  return nOut;
}

struct SdpParams { int wordSize, alignType, detailed, extendFront, sdpPrefix, recurse, noRecurseUnder, maxMatches; };

// results[job]: status, qPos, tPos, nBlocks; blocks of job at blocks[blockOff[job] ...), capacity blockOff[job + 1] - blockOff[job]
__global__ void __launch_bounds__(64) sdp_kernel(uint32_t nJobs, const uint8_t *q, const uint64_t *qOff, const uint8_t *t, const uint64_t *tOff,
                                                 Args a, SdpParams p, uint8_t *arena, size_t sliceBytes, uint32_t *counter,
                                                 bgpu_result *results, bgpu_block *blocks, const uint64_t *blockOff) {
  // one job per WARP, walked by lane 0: 32 jobs in one warp would take 32 different paths through this code and be
  // executed one after the other anyway (measured: 2,048 jobs on 64 warps 0.9 s, on 2,048 warps 32 lanes idle each far less)
  if (threadIdx.x & 31) return;
  Arena A;
  A.base = arena + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * sliceBytes; A.cap = sliceBytes;
  for (;;) {
    const uint32_t job = atomicAdd(counter, 1u);
    if (job >= nJobs) break;
    A.top = 0; A.oom = false;
    const uint8_t *qs = q + qOff[job], *ts = t + tOff[job];
    const uint32_t qLen = (uint32_t)(qOff[job + 1] - qOff[job]), tLen = (uint32_t)(tOff[job + 1] - tOff[job]);
    bgpu_result R{};
    R.blockOff = blockOff[job];
    bool bad = false;
    for (uint32_t i = 0; i < qLen; i++) if (base_code(qs[i]) > 4) bad = true;
    for (uint32_t i = 0; i < tLen; i++) if (base_code(ts[i]) > 4) bad = true;
    if (bad) { R.status = BGPU_JOB_BAD_INPUT; results[job] = R; continue; }
    Blk *out = reinterpret_cast<Blk *>(blocks + blockOff[job]);
    const uint32_t capOut = (uint32_t)(blockOff[job + 1] - blockOff[job]);
    Frame fr{qs, ts, qLen, tLen, p.wordSize, p.alignType, p.detailed, p.extendFront, p.sdpPrefix, p.recurse, p.noRecurseUnder, p.maxMatches};
    bool ovf = false;
    uint32_t n = sdp_align(A, a, fr, out, 0, capOut, ovf);
    if (A.oom || ovf || n > capOut) { R.status = BGPU_JOB_RANGE; results[job] = R; continue; }
    if (p.alignType == BGPU_LOCAL && n > 0) {                 // :604-612
      R.tPos = out[0].t; R.qPos = out[0].q;
      for (uint32_t i = 0; i < n; i++) { out[i].q -= R.qPos; out[i].t -= R.tPos; }
    }
    R.nBlocks = n;
    results[job] = R;
  }
}

}  // namespace sdp

// Host side: everything synchronous on `s`.  Returns a CUDA error code (0 = ok).
int run_sdp(const bgpu_scorefn *fn, const int *prm, float indelRate, int sdpIns, int sdpDel, uint32_t nJobs, const uint8_t *d_q,
            const uint64_t *d_qOff, const uint8_t *d_t, const uint64_t *d_tOff, uint8_t *d_arena, size_t sliceBytes, unsigned slices,
            uint32_t *d_counter, bgpu_result *d_results, bgpu_block *d_blocks, const uint64_t *d_blockOff, cudaStream_t s) {
  sdp::Args a;
  for (int i = 0; i < 25; i++) a.M[i] = fn->M[i];
  a.ins = fn->ins; a.del = fn->del; a.sdpIns = sdpIns; a.sdpDel = sdpDel; a.indelRate = indelRate;
  sdp::SdpParams p{prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], prm[7]};
  cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), s);
  const unsigned threads = 64, grid = (slices * 32 + threads - 1) / threads;       // `slices` warps
  sdp::sdp_kernel<<<grid, threads, 0, s>>>(nJobs, d_q, d_qOff, d_t, d_tOff, a, p, d_arena, sliceBytes, d_counter, d_results, d_blocks, d_blockOff);
  return (int)cudaGetLastError();
}

}  // namespace bgpu
