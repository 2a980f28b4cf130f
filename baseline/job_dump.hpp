// baseline/job_dump.hpp -- TEST / MEASUREMENT INFRASTRUCTURE (never part of the product).
//
// Included only by the instrumented twin of the reference program (baseline/_ref/blasrmc_dump, see
// baseline/Makefile + patch_blasr.py).  Writes every refinement job exactly as the reference's call sites
// hold it, so that the same job set can be replayed through the GPU library and through oracle/_ref:
//
//   kind 0 / 1   GuidedAlign / AffineGuidedAlign of RefineAlignment        alignment/Blasr.cpp:863-872
//   kind 4       AffineKBandAlign of AlignSubstring (-alignContigs gaps)   alignment/Blasr.cpp:1067-1076
//
// File format (little endian), one record per job, appended under a mutex (blasr's MapReads pthreads):
//   u32 magic 'BGJ1' | i32 kind | i32 band | u32 qLen | u32 tLen | u32 nBlocks | i32 ins, del, affineOpen, affineExtend
//   | i32 M[25] | i32 extra[5] (kind 4: hpInsOpen, hpInsExtend, insOpen, insExtend, del) | u8 q[qLen] | u8 t[tLen]
//   | u32 blocks[nBlocks][3] (qPos, tPos, length)
// The destination is $BGPU_DUMP; without it nothing is written.
#ifndef BGPU_JOB_DUMP_HPP_
#define BGPU_JOB_DUMP_HPP_
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <pthread.h>

namespace bgpu_dump {

struct Sink {
  FILE *f;
  pthread_mutex_t mu;
  Sink() : f(NULL) {
    pthread_mutex_init(&mu, NULL);
    const char *p = getenv("BGPU_DUMP");
    if (p && *p) f = fopen(p, "wb");
  }
  ~Sink() { if (f) fclose(f); }
};
inline Sink &sink() { static Sink s; return s; }

template <typename T_Blocks>
inline void Write(int kind, int band, const unsigned char *q, uint32_t qLen, const unsigned char *t, uint32_t tLen,
                  const T_Blocks *blocks, int ins, int del, int open, int ext, const int M[5][5], const int extra[5]) {
  Sink &s = sink();
  if (!s.f) return;
  pthread_mutex_lock(&s.mu);
  const uint32_t nB = blocks ? (uint32_t)blocks->size() : 0;
  uint32_t head[6] = {0x314a4742u, (uint32_t)kind, (uint32_t)band, qLen, tLen, nB};
  int32_t par[4 + 25 + 5];
  par[0] = ins; par[1] = del; par[2] = open; par[3] = ext;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) par[4 + i * 5 + j] = M[i][j];
  for (int i = 0; i < 5; i++) par[29 + i] = extra ? extra[i] : 0;
  fwrite(head, sizeof head, 1, s.f);
  fwrite(par, sizeof par, 1, s.f);
  if (qLen) fwrite(q, 1, qLen, s.f);
  if (tLen) fwrite(t, 1, tLen, s.f);
  for (uint32_t b = 0; b < nB; b++) {
    uint32_t v[3] = {(uint32_t)(*blocks)[b].qPos, (uint32_t)(*blocks)[b].tPos, (uint32_t)(*blocks)[b].length};
    fwrite(v, sizeof v, 1, s.f);
  }
  pthread_mutex_unlock(&s.mu);
}

// RefineAlignment: the slices and the guide handed to (Affine)GuidedAlign
template <typename T_Q, typename T_T, typename T_Cand, typename T_Fn>
inline void Guided(T_Q &q, T_T &t, T_Cand &cand, T_Fn &fn, int band, bool affine) {
  Write(affine ? 1 : 0, band, q.seq, q.length, t.seq, t.length, &cand.blocks, fn.ins, fn.del, fn.affineOpen, fn.affineExtend,
        fn.scoreMatrix, NULL);
}

struct NoBlocks { size_t size() const { return 0; } struct B { unsigned qPos, tPos, length; }; B operator[](size_t) const { return B(); } };

// AlignSubstring: AffineKBandAlign(q, t, matchMat, hpInsOpen, hpInsExtend, insOpen, insExtend, del, k, ..., Global)
template <typename T_Q, typename T_T>
inline void AffineKBand(T_Q &q, T_T &t, const int M[5][5], int hpInsOpen, int hpInsExtend, int insOpen, int insExtend, int del, int k) {
  const int extra[5] = {hpInsOpen, hpInsExtend, insOpen, insExtend, del};
  Write<NoBlocks>(4, k, q.seq, q.length, t.seq, t.length, NULL, 0, del, 0, 0, M, extra);
}

}  // namespace bgpu_dump
#endif
