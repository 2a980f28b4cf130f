/*
 * oracle/orc_anchor.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's suffix-array anchoring:
 *   MapReadToGenome                     common/algorithms/anchoring/MapBySuffixArray.h:209-309
 *   LocateAnchorBoundsInSuffixArray     common/algorithms/anchoring/MapBySuffixArray.h:24-207
 *   SuffixArray::StoreLCPBounds         common/datastructures/suffixarray/SuffixArray.h:928-1067
 *   SuffixArray::SearchLeftBound/Right  common/datastructures/suffixarray/SuffixArray.h:736-822
 *   DNATuple::FromStringLR              common/tuples/DNATuple.h:24-53
 * Pinned against the reference itself (oracle/_ref/libblasr_ref_anchor.so = oracle/ref_anchor.cpp) by
 * tests/test_anchor_oracle.py: the reference ships no test or golden vector for this path.
 *
 * Conventions shared with the device path (include/blasr_gpu.h, bgpu_map_reads): the genome buffer is readable one byte
 * past n (SuffixArray.h:1021 reads target[index[l] + lcpLength] before it checks the bound; blasr's genome always ends
 * with the 'N' FASTAReader.h:130 appends, so the byte is never reached there); removeEncompassedMatches reads its vectors
 * out of bounds in the reference (MapBySuffixArray.h:247-251) and is refused.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint8_t g3[256];
static int g3_ready = 0;
static void init3(void) {              /* ThreeBit, common/NucConversion.h:48-84 */
  if (g3_ready) return;
  memset(g3, 255, sizeof g3);
  const char *acgt = "ACGT";
  for (int i = 0; i < 4; i++) { g3[(unsigned char)acgt[i]] = (uint8_t)i; g3[(unsigned char)(acgt[i] + 32)] = (uint8_t)i; g3[i] = (uint8_t)i; }
  g3[4] = 4; g3['$'] = 5;
  const char *amb = "BDHKMNRSUVWY";
  for (const char *p = amb; *p; p++) { g3[(unsigned char)*p] = 4; g3[(unsigned char)(*p + 32)] = 4; }
  g3['x'] = 4; g3['_'] = 4;
  g3_ready = 1;
}
int orc_three_bit(int c) { init3(); return g3[c & 255]; }

typedef struct {
  const uint8_t *g; int64_t n; const uint32_t *sa;
  const uint32_t *startPos, *endPos; int prefixLen;
} sa_view;

/* SuffixArray.h:736-776 */
static int64_t left_bound(const sa_view *v, uint32_t off, uint8_t qc, int64_t l, int64_t r) {
  int64_t ll = l, lr = r;
  while (ll < lr) {
    const int64_t m = (ll + lr) / 2;
    const int64_t sufLen = v->n - (int64_t)v->sa[m];
    if (sufLen == (int64_t)off) { ll = m + 1; continue; }
    int comp;
    if (sufLen < (int64_t)off) comp = -1;
    else comp = (int)g3[v->g[(int64_t)v->sa[m] + off]] - (int)g3[qc];
    if (comp < 0) ll = m + 1; else lr = m;
  }
  return ll;
}
/* SuffixArray.h:778-816 */
static int64_t right_bound(const sa_view *v, uint32_t off, uint8_t qc, int64_t l, int64_t r) {
  int64_t rl = l, rr = r;
  while (rl < rr) {
    const int64_t m = (rl + rr) / 2;
    const int64_t sufLen = v->n - (int64_t)v->sa[m];
    if (sufLen == (int64_t)off) { rr = m; break; }
    if (sufLen < (int64_t)off) rr = m;
    else {
      const int comp = (int)g3[v->g[(int64_t)v->sa[m] + off]] - (int)g3[qc];
      if (comp <= 0) rl = m + 1; else rr = m;
    }
  }
  return rr;
}

typedef struct { uint32_t *lo, *hi; size_t n, cap; } bounds;
static void push(bounds *b, uint32_t lo, uint32_t hi) {
  if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 64; b->lo = realloc(b->lo, 4 * b->cap); b->hi = realloc(b->hi, 4 * b->cap); }
  b->lo[b->n] = lo; b->hi[b->n] = hi; b->n++;
}

/* SuffixArray.h:928-1067; returns lcpLength */
static uint32_t store_lcp_bounds(const sa_view *v, const uint8_t *query, uint32_t queryLength, int useLookupTable,
                                 int maxMatchLength, bounds *b, int stopOnceUnique) {
  int64_t l = 0, r = v->n;
  uint32_t lcp = 0;
  if (useLookupTable && v->startPos) {
    uint32_t tuple = 0;                       /* FromStringLR: OnlyACTG over tupleSize bases, then 2 bits per base */
    for (int i = 0; i < v->prefixLen; i++) if (g3[query[i]] > 3) return 0;
    for (int i = 0; i < v->prefixLen; i++) tuple = (tuple << 2) + g3[query[i]];
    l = v->startPos[tuple]; r = v->endPos[tuple];
    lcp = (uint32_t)v->prefixLen;
    if (l < r) push(b, (uint32_t)l, (uint32_t)r); else return 0;
  }
  while (l < r && lcp < queryLength) {
    if (stopOnceUnique && l == r - 1) break;
    if (maxMatchLength && lcp >= (uint32_t)maxMatchLength) break;
    if (g3[v->g[(int64_t)v->sa[l] + lcp]] >= 4) break;
    l = left_bound(v, lcp, query[lcp], l, r);
    r = right_bound(v, lcp, query[lcp], l, r);
    if (l == r || (int64_t)v->sa[l] + lcp >= v->n || g3[query[lcp]] >= 4 ||
        g3[v->g[(int64_t)v->sa[l] + lcp]] != g3[query[lcp]]) break;
    push(b, (uint32_t)l, (uint32_t)r);
    lcp++;
  }
  return lcp;
}

enum { P_MIN_PREFIX = 0, P_MIN_MATCH, P_EXPAND, P_USE_LOOKUP, P_MAX_ANCHORS, P_ADVANCE, P_MAX_LCP, P_STOP_UNIQUE, P_REMOVE_ENCOMPASSED };

/* One MapReadToGenome call; matches[3 * i ..] = (t, q, l); returns matchPosList.size() (entries beyond cap are counted, not
 * stored), -1 for the refused removeEncompassedMatches. */
int64_t orc_map_read(const uint8_t *genome, uint32_t n, const uint32_t *index, const uint32_t *startPos, const uint32_t *endPos,
                     int prefixLength, const uint8_t *read, uint32_t readLen, uint32_t subStart, uint32_t subEnd,
                     const int32_t *params, uint32_t *matches, uint64_t cap) {
  init3();
  if (params[P_REMOVE_ENCOMPASSED]) return -1;
  const uint32_t minPrefix = (uint32_t)params[P_MIN_PREFIX], minMatch = (uint32_t)params[P_MIN_MATCH];
  const int expand = params[P_EXPAND], advance = params[P_ADVANCE], maxLCP = params[P_MAX_LCP];
  sa_view v = {genome, (int64_t)n, index, startPos, endPos, prefixLength};
  /* MapBySuffixArray.h:219-222 */
  if (subEnd - subStart < minMatch) return 0;
  /* LocateAnchorBoundsInSuffixArray :39-42: nothing located, the three vectors stay empty */
  uint32_t nPos = 0;
  uint32_t *mLow = NULL, *mHigh = NULL, *mLen = NULL;
  if (!(minPrefix > 0 && subEnd - subStart < minPrefix)) {
    const uint32_t matchEnd = subEnd - minPrefix + 1;
    nPos = matchEnd - subStart;
    mLow = calloc(nPos ? nPos : 1, 4); mHigh = calloc(nPos ? nPos : 1, 4); mLen = calloc(nPos ? nPos : 1, 4);
    bounds b = {0, 0, 0, 0};
    uint32_t m = 0;
    for (uint32_t p = subStart; p < matchEnd; p++, m++) {
      b.n = 0;
      uint32_t lcp = store_lcp_bounds(&v, read + p, matchEnd - p, params[P_USE_LOOKUP], maxLCP, &b, params[P_STOP_UNIQUE]);
      mLow[m] = mHigh[m] = mLen[m] = 0;
      if (b.n > 0) {
        int s = (int)b.n;                                       /* lcpSearchLength :101-107 */
        while (s > 0 && b.lo[s - 1] == b.hi[s - 1]) { s--; lcp--; }
        mLow[m] = b.lo[s - 1]; mHigh[m] = b.hi[s - 1]; mLen[m] = minPrefix + s - 1;
        if (mLow[m] + 1 == mHigh[m]) {                          /* unique :134-174 */
          lcp = minPrefix + s - 1;
          int64_t refPos = (int64_t)index[mLow[m]] + lcp - 1, queryPos = (int64_t)p + lcp - 1;
          int extended = 0;
          while (refPos + 1 < (int64_t)n && queryPos + 1 < (int64_t)readLen && genome[refPos + 1] != 'N' &&
                 genome[refPos + 1] == read[queryPos + 1] && (maxLCP == 0 || lcp < (uint32_t)maxLCP)) {
            refPos++; queryPos++; lcp++; extended = 1;
          }
          if (extended) mLen[m] = lcp;
          else {
            if (s > 1) s = s - 1;
            mLow[m] = b.lo[s - 1]; mHigh[m] = b.hi[s - 1]; mLen[m] = minPrefix + s - 1;
          }
        } else {                                                /* not unique :176-195 */
          if (s > expand) s -= expand; else s = 1;
          mLow[m] = b.lo[s - 1]; mHigh[m] = b.hi[s - 1]; mLen[m] = minPrefix + s - 1;
        }
      }
      if (advance) {                                            /* :207-214 */
        int step = (int)lcp - advance; if (step < 1) step = 1;
        p += (uint32_t)step; m += (uint32_t)step;
      }
    }
    free(b.lo); free(b.hi);
  }
  /* MapBySuffixArray.h:266-305 */
  const uint32_t lookupPrefix = startPos ? (uint32_t)prefixLength : 0;        /* sa.lookupPrefixLength is 0 without a table */
  const uint32_t trim = (minMatch + 1 > lookupPrefix + 1) ? minMatch + 1 : lookupPrefix + 1;
  const uint32_t endOfMapping = subEnd < trim ? 0 : subEnd - trim;
  int64_t count = 0;
  for (uint32_t pos = subStart; pos < endOfMapping; pos++) {
    const uint32_t mi = pos - subStart;
    if (mi >= nPos) break;                    /* the reference asserts here (:279); callers keep minPrefix <= trim + 1 */
    if ((uint32_t)(mHigh[mi] - mLow[mi]) <= (uint32_t)params[P_MAX_ANCHORS]) {   /* DNALength arithmetic :280 */
      for (uint32_t mp = mLow[mi]; mp < mHigh[mi]; mp++) {
        if (mLen[mi] < minMatch) continue;
        if (mLen[mi] + pos > readLen) mLen[mi] = readLen - pos;
        if ((uint64_t)count < cap) { matches[3 * count] = index[mp]; matches[3 * count + 1] = pos; matches[3 * count + 2] = mLen[mi]; }
        count++;
      }
    }
  }
  free(mLow); free(mHigh); free(mLen);
  return count;
}
