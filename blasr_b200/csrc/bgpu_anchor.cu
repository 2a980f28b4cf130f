// bgpu_anchor.cu -- suffix-array anchoring on the device.
//
// What it replaces: MapReadToGenome (common/algorithms/anchoring/MapBySuffixArray.h:209-309) with
// LocateAnchorBoundsInSuffixArray (:24-207) and SuffixArray::StoreLCPBounds / SearchLeftBound / SearchRightBound
// (common/datastructures/suffixarray/SuffixArray.h:928-1067, 736-822), called per read and per strand at
// alignment/Blasr.cpp:2282-2296.
//
// Shape of the work: every read position p runs an independent longest-prefix search -- table look-up of the first
// lookupPrefixLength bases, then per base two binary searches over the current suffix-array interval, each probe a
// dependent pair of loads (index[m], then genome[index[m] + depth]).  Nothing here is arithmetic: the bound is memory
// latency, so the mapping is one THREAD per read position (tens of millions of independent probe chains in flight hide the
// latency; genome and index stay resident in HBM, a bacterial index sits in the 126 MB L2 whole).  The probe sequence is the
// reference's own (same midpoints, same early exits), so the bounds agree with it even on an index that is not perfectly
// sorted for the comparison in use.
//
//   locate_kernel   thread per (read, position): the interval [low, high) and match length LocateAnchorBounds stores
//   count_kernel    CTA per read: advanceExactMatches walk (sequential by definition, one thread), matches per read
//   scan_kernel     CSR offsets of the reads' match lists
//   emit_kernel     CTA per read: (t, q, l) triples in the reference's order (position ascending, suffix-array order inside)
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>
#include "bgpu_common.cuh"

namespace bgpu {

// An index entry on the device: the suffix's position and, packed 3 bits each, the ThreeBit codes of the CTX_BASES genome bases
// that follow the table's key (depths prefixLen .. prefixLen + 9 of the suffix; 6 = a byte outside the alphabet, 7 = past the
// end of the genome).  The binary searches of the depths right behind the key -- nearly all of them: a bacterial genome is
// unique after ~12 bases, a human one after ~17 -- then cost ONE 8-byte load per probe instead of the reference's dependent
// pair (index entry, then genome byte).  Built on the device by bgpu_set_suffix_array; 8 B per base.
constexpr uint32_t CTX_BASES = 10;
struct AnchorIndex { uint2 *ent = nullptr; uint32_t *startT = nullptr, *endT = nullptr; uint64_t n = 0, refGen = 0; uint32_t prefixLen = 0; };

struct MapArgs {
  const uint8_t *g; uint64_t n;
  const uint2 *ent; const uint32_t *startT, *endT; uint32_t prefixLen;      // startT == nullptr: no table (entries' context starts at depth 0)
  const uint8_t *reads; const uint64_t *readOff; const uint32_t *subS, *subE;
  const uint64_t *posOff; uint32_t nReads; uint64_t totalPos;
  bgpu_anchor_params p;
  uint32_t *lo, *hi, *len, *lcp;                                // per searched position (lcp only with advanceExactMatches)
  uint64_t *counts, *matchOff;
  bgpu_match *matches;
};

__device__ __forceinline__ int code7(int c) { return c == 255 ? 6 : c; }     // ThreeBit order kept, the outsider next to the alphabet

// What the searches compare at depth `off` of the suffix behind index entry m: its ThreeBit code (code7), -1 when the suffix
// ends exactly there, -2 when it is shorter still.  *at = the suffix's position.
__device__ __forceinline__ int probe(const MapArgs &a, int64_t m, uint32_t off, uint64_t *at) {
  const uint2 e = __ldg(a.ent + m);
  *at = e.x;
  const uint32_t rel = off - a.prefixLen;
  if (rel < CTX_BASES) {
    const int c = (int)(e.y >> (3 * rel)) & 7;
    if (c != 7) return c;
    return (a.n - (uint64_t)e.x == (uint64_t)off) ? -1 : -2;
  }
  const int64_t sufLen = (int64_t)a.n - (int64_t)e.x;
  if (sufLen == (int64_t)off) return -1;
  if (sufLen < (int64_t)off) return -2;
  return code7(base_code(__ldg(a.g + (uint64_t)e.x + off)));
}

// SuffixArray::SearchLeftBound, SuffixArray.h:736-776 (a suffix that has ended sorts before every base)
__device__ __forceinline__ int64_t left_bound(const MapArgs &a, uint32_t off, int qc7, int64_t l, int64_t r) {
  int64_t ll = l, lr = r;
  uint64_t at;
  while (ll < lr) {
    const int64_t m = (ll + lr) / 2;
    if (probe(a, m, off, &at) < qc7) ll = m + 1; else lr = m;
  }
  return ll;
}
// SuffixArray::SearchRightBound, SuffixArray.h:778-816 (meeting a suffix that ends exactly at this depth ends the search, :786-789)
__device__ __forceinline__ int64_t right_bound(const MapArgs &a, uint32_t off, int qc7, int64_t l, int64_t r) {
  int64_t rl = l, rr = r;
  uint64_t at;
  while (rl < rr) {
    const int64_t m = (rl + rr) / 2;
    const int k = probe(a, m, off, &at);
    if (k == -1) { rr = m; break; }
    if (k != -2 && k <= qc7) rl = m + 1; else rr = m;
  }
  return rr;
}

constexpr int RING = 16;          // bounds of the last RING depths (expand <= RING - 2)

__global__ void __launch_bounds__(256) locate_kernel(const MapArgs a) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.totalPos) return;
  uint32_t r0 = 0, r1 = a.nReads;                   // read of this position: last r with posOff[r] <= gid
  while (r1 - r0 > 1) { const uint32_t m = (r0 + r1) / 2; if (__ldg(a.posOff + m) <= gid) r0 = m; else r1 = m; }
  const uint32_t rd = r0;
  const uint32_t mIdx = (uint32_t)(gid - a.posOff[rd]);
  const uint64_t rOff = a.readOff[rd];
  const uint32_t readLen = (uint32_t)(a.readOff[rd + 1] - rOff);
  const uint8_t *read = a.reads + rOff;
  const uint32_t minPrefix = a.p.minPrefixMatchLength;
  const uint32_t p = a.subS[rd] + mIdx;
  const uint32_t matchEnd = a.subE[rd] - minPrefix + 1;
  const uint32_t queryLength = matchEnd - p;        // MapBySuffixArray.h:66: the search may not run past matchEnd
  const uint8_t *query = read + p;
  const int expand = a.p.expand;
  const uint32_t maxLCP = (uint32_t)a.p.maxLCPLength;

  // ---- StoreLCPBounds, SuffixArray.h:928-1067
  int64_t l = 0, r = (int64_t)a.n;
  uint32_t lcp = 0, S = 0;                          // S = lcpLeftBounds.size()
  uint32_t lo0 = 0, hi0 = 0, loC = 0, hiC = 0, loP = 0, hiP = 0;     // bounds at depth 0, S - 1, S - 2
  uint32_t ringLo[RING], ringHi[RING];
  const bool useRing = expand > 1;
  bool live = true;
  auto push = [&](uint32_t lo, uint32_t hi) {
    if (S == 0) { lo0 = lo; hi0 = hi; }
    loP = loC; hiP = hiC; loC = lo; hiC = hi;
    if (useRing) { ringLo[S & (RING - 1)] = lo; ringHi[S & (RING - 1)] = hi; }
    S++;
  };
  if (a.p.useLookupTable && a.startT) {
    uint32_t tuple = 0; bool ok = true;             // DNATuple::FromStringLR, tuples/DNATuple.h:24-53
    for (uint32_t i = 0; i < a.prefixLen; i++) { const int c = base_code(__ldg(query + i)); ok = ok && c <= 3; tuple = (tuple << 2) + (uint32_t)(c & 3); }
    if (!ok) live = false;
    else {
      l = __ldg(a.startT + tuple); r = __ldg(a.endT + tuple);
      if (l < r) { lcp = a.prefixLen; push((uint32_t)l, (uint32_t)r); } else live = false;     // "return 0": lcpLength 0
    }
  }
  if (live) {
    while (l < r && lcp < queryLength) {
      if (a.p.stopMappingOnceUnique && l == r - 1) break;
      if (maxLCP && lcp >= maxLCP) break;
      uint64_t at;
      const int k0 = probe(a, l, lcp, &at);          // SuffixArray.h:1021: the genome at and beyond n reads as 'N'
      if (k0 < 0 || k0 >= 4) break;
      const int qc = base_code(__ldg(query + lcp)), qc7 = code7(qc);
      l = left_bound(a, lcp, qc7, l, r);
      r = right_bound(a, lcp, qc7, l, r);
      if (l == r) break;
      const int k1 = probe(a, l, lcp, &at);
      if (k1 < 0 || qc >= 4 || k1 != qc7) break;     // :1040-1047
      push((uint32_t)l, (uint32_t)r);
      lcp++;
    }
  }

  // ---- LocateAnchorBoundsInSuffixArray, MapBySuffixArray.h:90-214
  uint32_t mLow = 0, mHigh = 0, mLen = 0;
  if (S > 0) {
    auto at = [&](uint32_t s, uint32_t &lo, uint32_t &hi) {        // lowMatchBound[s - 1], highMatchBound[s - 1]
      if (s == S) { lo = loC; hi = hiC; }
      else if (s == 1) { lo = lo0; hi = hi0; }
      else if (s + 1 == S) { lo = loP; hi = hiP; }
      else { lo = ringLo[(s - 1) & (RING - 1)]; hi = ringHi[(s - 1) & (RING - 1)]; }
    };
    uint32_t s = S;            // every stored interval is non-empty (SuffixArray.h:1043), so the shrink loop :104-107 never runs
    mLow = loC; mHigh = hiC; mLen = minPrefix + s - 1;
    if (mLow + 1 == mHigh) {                                        // unique: extend along the genome :134-174
      lcp = minPrefix + s - 1;
      int64_t refPos = (int64_t)__ldg(&a.ent[mLow].x) + lcp - 1, queryPos = (int64_t)p + lcp - 1;
      bool extended = false;
      while (refPos + 1 < (int64_t)a.n && queryPos + 1 < (int64_t)readLen) {
        const uint8_t gc = __ldg(a.g + refPos + 1);
        if (gc == 'N' || gc != __ldg(read + queryPos + 1) || !(maxLCP == 0 || lcp < maxLCP)) break;
        refPos++; queryPos++; lcp++; extended = true;
      }
      if (extended) mLen = lcp;
      else {
        if (s > 1) s--;
        at(s, mLow, mHigh); mLen = minPrefix + s - 1;
      }
    } else {                                                        // not unique: back off by `expand` depths :176-195
      if ((int)s > expand) s -= (uint32_t)expand; else s = 1;
      at(s, mLow, mHigh); mLen = minPrefix + s - 1;
    }
  }
  a.lo[gid] = mLow; a.hi[gid] = mHigh; a.len[gid] = mLen;
  if (a.lcp) a.lcp[gid] = lcp;
}

// matches position mi of a read contributes, and the length they carry (MapBySuffixArray.h:276-305, including the trim at
// :292-299 that, once applied, is tested against minMatchLength again for the following suffixes)
__device__ __forceinline__ uint32_t match_count(const MapArgs &a, uint32_t lo, uint32_t hi, uint32_t len, uint32_t pos, uint32_t readLen,
                                                uint32_t &outLen) {
  outLen = len;
  if ((uint32_t)(hi - lo) > (uint32_t)a.p.maxAnchorsPerPosition || hi <= lo || len < a.p.minMatchLength) return 0;
  if (len + pos > readLen) { outLen = readLen - pos; if (outLen < a.p.minMatchLength) return 1; }
  return hi - lo;
}

__device__ __forceinline__ uint32_t end_of_mapping(const MapArgs &a, uint32_t subEnd) {
  const uint32_t lookupPrefix = a.startT ? a.prefixLen : 0;
  const uint32_t trim = max(a.p.minMatchLength + 1, lookupPrefix + 1);
  return subEnd < trim ? 0 : subEnd - trim;
}

__global__ void __launch_bounds__(256) count_kernel(const MapArgs a) {
  const uint32_t rd = blockIdx.x;
  const uint64_t base = a.posOff[rd];
  const uint32_t nPos = (uint32_t)(a.posOff[rd + 1] - base);
  const uint32_t subS = a.subS[rd], readLen = (uint32_t)(a.readOff[rd + 1] - a.readOff[rd]);
  if (a.p.advanceExactMatches && nPos) {            // :207-214: the positions visited depend on the lcp of the ones before
    if (threadIdx.x == 0) {
      uint32_t m = 0;
      while (m < nPos) {
        int step = (int)a.lcp[base + m] - a.p.advanceExactMatches; if (step < 1) step = 1;
        const uint32_t next = m + 1 + (uint32_t)step;
        for (uint32_t k = m + 1; k < next && k < nPos; k++) { a.lo[base + k] = 0; a.hi[base + k] = 0; a.len[base + k] = 0; }
        m = next;
      }
    }
    __syncthreads();
  }
  const uint32_t eom = end_of_mapping(a, a.subE[rd]);
  uint64_t cnt = 0;
  for (uint32_t mi = threadIdx.x; mi < nPos; mi += blockDim.x) {
    const uint32_t pos = subS + mi;
    if (pos >= eom) break;
    uint32_t L;
    cnt += match_count(a, a.lo[base + mi], a.hi[base + mi], a.len[base + mi], pos, readLen, L);
  }
  __shared__ uint64_t sh[8];
  for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t s = 0; for (int w = 0; w < 8; w++) s += sh[w]; a.counts[rd] = s; }
}

__global__ void __launch_bounds__(1024) scan_kernel(const uint64_t *counts, uint64_t *off, uint32_t n) {
  __shared__ uint64_t part[1024];
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
  uint64_t s = 0;
  for (uint32_t i = b; i < e; i++) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t run = 0; for (int i = 0; i < 1024; i++) { const uint64_t v = part[i]; part[i] = run; run += v; } off[n] = run; }
  __syncthreads();
  uint64_t run = part[threadIdx.x];
  for (uint32_t i = b; i < e; i++) { off[i] = run; run += counts[i]; }
}

__global__ void __launch_bounds__(256) emit_kernel(const MapArgs a) {
  const uint32_t rd = blockIdx.x;
  const uint64_t base = a.posOff[rd];
  const uint32_t nPos = (uint32_t)(a.posOff[rd + 1] - base);
  const uint32_t subS = a.subS[rd], readLen = (uint32_t)(a.readOff[rd + 1] - a.readOff[rd]);
  const uint32_t eom = end_of_mapping(a, a.subE[rd]);
  const uint32_t nLive = eom > subS ? min(nPos, eom - subS) : 0;
  __shared__ uint64_t warpSum[8];
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) carry = a.matchOff[rd];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t c0 = 0; c0 < nLive; c0 += blockDim.x) {
    const uint32_t mi = c0 + threadIdx.x;
    uint32_t lo = 0, L = 0, cnt = 0, pos = subS + mi;
    if (mi < nLive) { lo = a.lo[base + mi]; cnt = match_count(a, lo, a.hi[base + mi], a.len[base + mi], pos, readLen, L); }
    uint64_t inc = cnt;                             // inclusive scan over the CTA
    for (int o = 1; o < 32; o <<= 1) { const uint64_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) warpSum[wid] = inc;
    __syncthreads();
    uint64_t before = carry;
    for (int w = 0; w < wid; w++) before += warpSum[w];
    uint64_t at = before + inc - cnt;
    for (uint32_t k = 0; k < cnt; k++) { bgpu_match mt; mt.t = __ldg(&a.ent[lo + k].x); mt.q = pos; mt.l = L; a.matches[at + k] = mt; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + inc;
    __syncthreads();
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------

struct AnchorState {                // per context: device / pinned buffers of the last bgpu_map_reads, kept for the next call
  void *dev = nullptr; size_t devBytes = 0;
  void *pin = nullptr; size_t pinBytes = 0;
  bgpu_match *dMatches = nullptr; size_t matchCap = 0;
  MapArgs args{}; bool valid = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double ms[2] = {0, 0}; uint64_t h2d = 0, d2h = 0, total = 0;
};

static AnchorIndex g_index[64];

__global__ void __launch_bounds__(256) build_entries_kernel(const uint32_t *sa, const uint8_t *g, uint64_t n, uint32_t prefixLen, uint2 *ent) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t at = sa[i];
  uint32_t w = 0;
  for (uint32_t k = 0; k < CTX_BASES; k++) {
    const uint64_t pos = at + prefixLen + k;
    const uint32_t c = pos < n ? (uint32_t)code7(base_code(g[pos])) : 7u;
    w |= c << (3 * k);
  }
  ent[i] = make_uint2((uint32_t)at, w);
}

#define ACK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[256]; snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); err = b_; return e_ == cudaErrorMemoryAllocation ? BGPU_E_OOM : BGPU_E_CUDA; } } while (0)

int anchor_set_index(int device, const uint8_t *genome, uint64_t gN, uint64_t refGen, const uint32_t *index, uint64_t n, const uint32_t *startT,
                     const uint32_t *endT, uint32_t prefixLen, std::string &err) {
  if (device < 0 || device >= 64) return BGPU_E_INVALID;
  AnchorIndex &ix = g_index[device];
  ACK(cudaDeviceSynchronize());
  cudaFree(ix.ent); cudaFree(ix.startT); cudaFree(ix.endT);
  ix = AnchorIndex();
  if (!n) return BGPU_OK;
  if (!index || n >= 0xFFFFFFFFull) { err = "bgpu_set_suffix_array: index is NULL or the genome does not fit SAIndex (uint32)"; return BGPU_E_INVALID; }
  if ((startT == nullptr) != (endT == nullptr) || (startT && (prefixLen < 1 || prefixLen > 14))) {
    err = "bgpu_set_suffix_array: startPosTable and endPosTable come together, lookupPrefixLength 1..14"; return BGPU_E_INVALID;
  }
  if (!genome || gN != n) { err = "bgpu_set_suffix_array: call bgpu_set_reference first, with the genome this array indexes (same length)"; return BGPU_E_INVALID; }
  if (startT) ix.prefixLen = prefixLen;
  uint32_t *dSa = nullptr;
  ACK(cudaMalloc(&dSa, sizeof(uint32_t) * (n + 4)));
  cudaError_t e = cudaMemcpy(dSa, index, sizeof(uint32_t) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&ix.ent, sizeof(uint2) * (n + 4));
  if (e == cudaSuccess) { build_entries_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dSa, genome, n, ix.prefixLen, ix.ent); e = cudaDeviceSynchronize(); }
  cudaFree(dSa);
  if (e != cudaSuccess) { cudaFree(ix.ent); ix = AnchorIndex(); ACK(e); }
  if (startT) {
    const size_t tl = (size_t)1 << (2 * prefixLen);
    ACK(cudaMalloc(&ix.startT, sizeof(uint32_t) * tl)); ACK(cudaMalloc(&ix.endT, sizeof(uint32_t) * tl));
    ACK(cudaMemcpy(ix.startT, startT, sizeof(uint32_t) * tl, cudaMemcpyHostToDevice));
    ACK(cudaMemcpy(ix.endT, endT, sizeof(uint32_t) * tl, cudaMemcpyHostToDevice));
  }
  ix.n = n; ix.refGen = refGen;
  return BGPU_OK;
}

// SuffixArray::BuildLookupTable (SuffixArray.h:193-250), what blasr runs when the .sa file carries no table (Blasr.cpp:4419)
// and sawriter before it writes one (SAWriter.cpp:225): for every lookupPrefixLength-mer the run of suffix-array entries that
// start with it.  Host code, one sequential pass in suffix-array order, with the reference's own edge behaviour: the last
// prefixLength - 1 ENTRIES of the array are never looked at, a suffix whose k-mer ends exactly at the end of the text closes the
// current run, and a k-mer that cannot be coded (N, or the text ending inside it: bytes at and beyond n read as 'N') leaves
// the start of the tuple seen before it rewritten.  the CPU tests pin it against the reference's own tables.
int build_lookup_table(const uint8_t *g, uint64_t n, const uint32_t *index, uint32_t L, uint32_t *startT, uint32_t *endT) {
  if (!g || !index || !startT || !endT || L < 1 || L > 14 || n >= 0xFFFFFFFFull) return BGPU_E_INVALID;
  const uint64_t tableLen = (uint64_t)1 << (2 * L);
  for (uint64_t i = 0; i < tableLen; i++) startT[i] = endT[i] = 0;
  if (n < L) return BGPU_OK;
  auto tupleAt = [&](uint64_t at, uint64_t &tuple) -> bool {           // DNATuple::FromStringLR, DNATuple.h:24-53
    uint64_t v = 0;
    for (uint32_t i = 0; i < L; i++) {
      const int c = at + i < n ? base_code(g[at + i]) : 4;
      if (c > 3) return false;
      v = (v << 2) + (uint64_t)c;
    }
    tuple = v;
    return true;
  };
  const uint64_t lim = n - L + 1;
  uint64_t pos = 0, cur = 0;
  do {
    while (pos < lim && (uint64_t)index[pos] + L > n) pos++;
    if (pos >= lim) break;
    while (pos < lim && !tupleAt(index[pos], cur)) ++pos;
    startT[cur] = (uint32_t)pos;
    pos++;
    while (pos < lim && (uint64_t)index[pos] + L < n) {
      uint64_t next = 0;
      tupleAt(index[pos], next);                                        // a k-mer that cannot be coded compares as tuple 0
      if (next != cur) break;
      pos++;
    }
    endT[cur] = (uint32_t)pos;
  } while (pos < lim && cur + 1 < tableLen);
  return BGPU_OK;
}

void anchor_free_state(AnchorState *st) {
  if (!st) return;
  cudaFree(st->dev); cudaFree(st->dMatches); if (st->pin) cudaFreeHost(st->pin);
  for (auto &e : st->ev) if (e) cudaEventDestroy(e);
  delete st;
}

static int run_kernels(AnchorState *st, cudaStream_t s, std::string &err, bool emitOnly = false) {
  MapArgs &a = st->args;
  if (!emitOnly) {
    ACK(cudaEventRecord(st->ev[0], s));
    if (a.totalPos) locate_kernel<<<(unsigned)((a.totalPos + 255) / 256), 256, 0, s>>>(a);
    ACK(cudaEventRecord(st->ev[1], s));
    count_kernel<<<a.nReads, 256, 0, s>>>(a);
    scan_kernel<<<1, 1024, 0, s>>>(a.counts, a.matchOff, a.nReads);
  } else {
    emit_kernel<<<a.nReads, 256, 0, s>>>(a);
    ACK(cudaEventRecord(st->ev[2], s));
  }
  ACK(cudaGetLastError());
  return BGPU_OK;
}

int anchor_map(int device, const uint8_t *genome, uint64_t gN, uint64_t refGen, cudaStream_t s, AnchorState **stp, const bgpu_anchor_params *p,
               const uint8_t *reads, const uint64_t *readOff, uint32_t nReads, const uint32_t *subS, const uint32_t *subE,
               uint64_t *matchOff, const bgpu_match **matches, std::string &err) {
  const AnchorIndex &ix = g_index[device];
  if (!ix.ent) { err = "bgpu_map_reads without bgpu_set_suffix_array on this device"; return BGPU_E_INVALID; }
  if (!genome || gN != ix.n || refGen != ix.refGen) { err = "bgpu_map_reads: the genome was replaced after bgpu_set_suffix_array (the index entries carry its bases): set the reference, then the suffix array"; return BGPU_E_INVALID; }
  if (p->removeEncompassedMatches) { err = "removeEncompassedMatches reads out of bounds in the reference (MapBySuffixArray.h:247-251)"; return BGPU_E_INVALID; }
  if (p->expand < 0 || p->expand > RING - 2 || p->maxLCPLength < 0 || p->advanceExactMatches < 0 || p->maxAnchorsPerPosition < 0) {
    err = "bgpu_map_reads: expand 0..14, maxLCPLength / advanceExactMatches / maxAnchorsPerPosition >= 0"; return BGPU_E_INVALID;
  }
  const bool table = p->useLookupTable && ix.startT;
  const uint32_t lookupPrefix = ix.startT ? ix.prefixLen : 0;
  if (table && ix.prefixLen > p->minPrefixMatchLength) { err = "lookupPrefixLength > minPrefixMatchLength: the reference reads the tuple past the subread"; return BGPU_E_INVALID; }
  if (p->minPrefixMatchLength > std::max(p->minMatchLength, lookupPrefix) + 2) { err = "minPrefixMatchLength > max(minMatchLength, lookupPrefixLength) + 2: the reference asserts (MapBySuffixArray.h:279)"; return BGPU_E_INVALID; }
  if (!*stp) { *stp = new AnchorState(); for (auto &e : (*stp)->ev) ACK(cudaEventCreate(&e)); }
  AnchorState *st = *stp;
  st->valid = false;
  matchOff[0] = 0;
  *matches = nullptr;
  if (nReads == 0) return BGPU_OK;
  // searched positions per read: numSearchedPositions of LocateAnchorBounds (:45-46), 0 where MapReadToGenome (:219-222) or
  // LocateAnchorBounds (:39-42) return before searching
  const uint64_t totR = readOff[nReads];
  std::vector<uint64_t> posOff((size_t)nReads + 1, 0);
  std::vector<uint32_t> hs(nReads), he(nReads);
  for (uint32_t i = 0; i < nReads; i++) {
    const uint64_t len = readOff[i + 1] - readOff[i];
    if (len >= 0xFFFFFFFFull) { err = "read longer than DNALength"; return BGPU_E_INVALID; }
    const uint32_t b = subS ? subS[i] : 0, e = subE ? subE[i] : (uint32_t)len;
    if (b > e || e > len) { err = "subread outside the read"; return BGPU_E_INVALID; }
    hs[i] = b; he[i] = e;
    uint32_t np = 0;
    if (e - b >= p->minMatchLength && !(p->minPrefixMatchLength > 0 && e - b < p->minPrefixMatchLength)) np = e - p->minPrefixMatchLength + 1 - b;
    posOff[i + 1] = posOff[i] + np;
  }
  const uint64_t totalPos = posOff[nReads];
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t szReads = up(totR + 16), szOff = up(8 * ((size_t)nReads + 1)), szSub = up(4 * (size_t)nReads), szPos = up(4 * (totalPos + 1));
  const size_t need = szReads + 3 * szOff + 2 * szSub + (p->advanceExactMatches ? 4 : 3) * szPos + up(8 * (size_t)nReads);
  if (st->devBytes < need) {
    cudaFree(st->dev); st->dev = nullptr; st->devBytes = 0;
    ACK(cudaMalloc(&st->dev, need));
    st->devBytes = need;
  }
  uint8_t *d = (uint8_t *)st->dev;
  auto take = [&](size_t b) { uint8_t *q = d; d += b; return q; };
  MapArgs &a = st->args;
  a = MapArgs();
  a.g = genome; a.n = ix.n; a.ent = ix.ent; a.startT = ix.startT; a.endT = ix.endT; a.prefixLen = ix.prefixLen;
  a.nReads = nReads; a.totalPos = totalPos; a.p = *p;
  uint8_t *dReads = take(szReads); uint64_t *dReadOff = (uint64_t *)take(szOff), *dPosOff = (uint64_t *)take(szOff), *dMatchOff = (uint64_t *)take(szOff);
  uint32_t *dS = (uint32_t *)take(szSub), *dE = (uint32_t *)take(szSub);
  a.lo = (uint32_t *)take(szPos); a.hi = (uint32_t *)take(szPos); a.len = (uint32_t *)take(szPos);
  a.lcp = p->advanceExactMatches ? (uint32_t *)take(szPos) : nullptr;
  a.counts = (uint64_t *)take(up(8 * (size_t)nReads));
  a.reads = dReads; a.readOff = dReadOff; a.posOff = dPosOff; a.matchOff = dMatchOff; a.subS = dS; a.subE = dE;
  ACK(cudaMemcpyAsync(dReads, reads, totR, cudaMemcpyHostToDevice, s));
  ACK(cudaMemcpyAsync(dReadOff, readOff, 8 * ((size_t)nReads + 1), cudaMemcpyHostToDevice, s));
  ACK(cudaMemcpyAsync(dPosOff, posOff.data(), 8 * ((size_t)nReads + 1), cudaMemcpyHostToDevice, s));
  ACK(cudaMemcpyAsync(dS, hs.data(), 4 * (size_t)nReads, cudaMemcpyHostToDevice, s));
  ACK(cudaMemcpyAsync(dE, he.data(), 4 * (size_t)nReads, cudaMemcpyHostToDevice, s));
  st->h2d = totR + 24 * ((size_t)nReads + 1);
  int rc = run_kernels(st, s, err);
  if (rc) return rc;
  ACK(cudaMemcpyAsync(matchOff, dMatchOff, 8 * ((size_t)nReads + 1), cudaMemcpyDeviceToHost, s));
  ACK(cudaStreamSynchronize(s));
  const uint64_t total = matchOff[nReads];
  st->total = total;
  if (st->matchCap < total + 1) {
    cudaFree(st->dMatches); st->dMatches = nullptr; st->matchCap = 0;
    ACK(cudaMalloc(&st->dMatches, sizeof(bgpu_match) * (total + 1 + total / 8)));
    st->matchCap = total + 1 + total / 8;
  }
  if (st->pinBytes < sizeof(bgpu_match) * (total + 1)) {
    if (st->pin) cudaFreeHost(st->pin);
    st->pin = nullptr; st->pinBytes = 0;
    ACK(cudaHostAlloc(&st->pin, sizeof(bgpu_match) * (total + 1 + total / 8), cudaHostAllocDefault));
    st->pinBytes = sizeof(bgpu_match) * (total + 1 + total / 8);
  }
  a.matches = st->dMatches;
  rc = run_kernels(st, s, err, true);
  if (rc) return rc;
  ACK(cudaMemcpyAsync(st->pin, st->dMatches, sizeof(bgpu_match) * total, cudaMemcpyDeviceToHost, s));
  ACK(cudaStreamSynchronize(s));
  st->d2h = 8 * ((size_t)nReads + 1) + sizeof(bgpu_match) * total;
  float f0 = 0, f1 = 0;
  cudaEventElapsedTime(&f0, st->ev[0], st->ev[1]); cudaEventElapsedTime(&f1, st->ev[1], st->ev[2]);
  st->ms[0] = f0; st->ms[1] = f1;
  st->valid = true;
  *matches = (const bgpu_match *)st->pin;
  return BGPU_OK;
}

int anchor_rerun(AnchorState *st, cudaStream_t s, std::string &err) {
  if (!st || !st->valid) { err = "bgpu_map_rerun: no bgpu_map_reads to repeat on this context"; return BGPU_E_BUSY; }
  int rc = run_kernels(st, s, err);
  if (rc) return rc;
  rc = run_kernels(st, s, err, true);
  if (rc) return rc;
  ACK(cudaStreamSynchronize(s));
  float f0 = 0, f1 = 0;
  cudaEventElapsedTime(&f0, st->ev[0], st->ev[1]); cudaEventElapsedTime(&f1, st->ev[1], st->ev[2]);
  st->ms[0] = f0; st->ms[1] = f1;
  return BGPU_OK;
}

int anchor_timing(const AnchorState *st, double ms[2], uint64_t *positions, uint64_t *h2d, uint64_t *d2h) {
  if (!st || !st->valid) return BGPU_E_BUSY;
  if (ms) { ms[0] = st->ms[0]; ms[1] = st->ms[1]; }
  if (positions) *positions = st->args.totalPos;
  if (h2d) *h2d = st->h2d;
  if (d2h) *d2h = st->d2h;
  return BGPU_OK;
}

}  // namespace bgpu
