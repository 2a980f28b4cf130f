// baseline/gpu_refine.hpp -- the patch of INTEGRATION.md section 2, compiled for real into baseline/_ref/blasrmc_gpu.
//
// Included by the patched copy of alignment/Blasr.cpp right above RefineAlignments (Blasr.cpp:2163), whose first
// statement becomes `if (BgpuRefineAlignments(...)) return;`.  Everything else of the reference program -- anchoring,
// SDPAlign, filters, mapQV, printers, the pthread driver -- is the reference's own code, unmodified.
//
// What it replaces, for the default path (useGuidedAlign, not -global):
//   RefineAlignment: slices (Blasr.cpp:850-859), AffineGuidedAlign / GuidedAlign (:862-873), ComputeAlignmentStats
//   (:875-878), copy-back (:888-914); RefineAlignments' sort (:2178-2180).
// Each MapReads pthread calls this with the candidates of one read; blasr_gpu::RefineService merges the concurrent
// calls of all -nproc threads into one GPU ticket.
#ifndef BGPU_GPU_REFINE_HPP_
#define BGPU_GPU_REFINE_HPP_
#include <cstdlib>
#include "blasr_gpu_adapter.hpp"

static blasr_gpu::RefineService &BgpuService(int nProc) {
  // MapReads runs more pthreads than the host has cores once the GPU takes the refinement: waiting threads must sleep
  static const int once = setenv("BGPU_BLOCKING_SYNC", "1", 0);
  (void)once;
  static blasr_gpu::RefineService svc(getenv("BGPU_DEVICE") ? atoi(getenv("BGPU_DEVICE")) : 0, nProc,
                                      getenv("BGPU_BATCH_WAIT_US") ? atoi(getenv("BGPU_BATCH_WAIT_US")) : 300,
                                      getenv("BGPU_SERVICE_CONTEXTS") ? atoi(getenv("BGPU_SERVICE_CONTEXTS")) : 3);
  return svc;
}

template<typename T_RefSequence, typename T_Sequence>
bool BgpuRefineAlignments(vector<T_Sequence*> &bothQueryStrands, T_RefSequence &genome,
                          vector<T_AlignmentCandidate*> &alignmentPtrs, MappingParameters &params,
                          MappingBuffers &mappingBuffers) {
  if (params.doGlobalAlignment || !params.useGuidedAlign) return false;     // the other branches stay the reference's
  DistanceMatrixScoreFunction<DNASequence, FASTQSequence> distScoreFn;
  params.InitializeScoreFunction(distScoreFn);
  distScoreFn.InitializeScoreMatrix(SMRTDistanceMatrix);

  T_Sequence &query = *bothQueryStrands[0];
  blasr_gpu::RefineBatch batch;
  vector<DNASequence> tSeqs(alignmentPtrs.size());
  vector<FASTQSequence> qSeqs(alignmentPtrs.size());
  for (UInt i = 0; i < alignmentPtrs.size(); i++) {
    T_AlignmentCandidate &c = *alignmentPtrs[i];
    if (c.blocks.size() == 0) continue;
    int lastBlock = c.blocks.size() - 1;
    tSeqs[i].Copy(c.tAlignedSeq, c.tPos, c.blocks[lastBlock].tPos + c.blocks[lastBlock].length);
    qSeqs[i].ReferenceSubstring(query, c.qAlignedSeqPos + c.qPos, c.blocks[lastBlock].qPos + c.blocks[lastBlock].length);
    batch.Add(qSeqs[i].seq, qSeqs[i].length, tSeqs[i].seq, tSeqs[i].length, c.blocks);
  }
  if (batch.size() > 0)
    BgpuService(params.nProc).Run(batch, distScoreFn, params.affineAlign ? params.bandSize : params.guidedAlignBandSize,
                                  params.affineAlign);
  UInt j = 0;
  for (UInt i = 0; i < alignmentPtrs.size(); i++) {
    T_AlignmentCandidate &c = *alignmentPtrs[i];
    if (c.blocks.size() == 0) continue;
    T_AlignmentCandidate refinedAlignment;
    batch.Store(j++, refinedAlignment);
    c.blocks.clear();
    c.blocks = refinedAlignment.blocks;
    c.CopyStats(refinedAlignment);
    c.gaps = refinedAlignment.gaps;
    c.score = refinedAlignment.score;
    c.nCells = refinedAlignment.nCells;
    c.tAlignedSeq.TakeOwnership(tSeqs[i]);
    c.ReassignQSequence(qSeqs[i]);
    c.tAlignedSeqPos += c.tPos;
    c.qAlignedSeqPos += c.qPos;
    c.tPos = refinedAlignment.tPos;
    c.qPos = refinedAlignment.qPos;
    c.tAlignedSeqLength = tSeqs[i].length;
    c.qAlignedSeqLength = qSeqs[i].length;
  }
  if (params.sortRefinedAlignments)
    std::sort(alignmentPtrs.begin(), alignmentPtrs.end(), SortAlignmentPointersByScore());
  return true;
}
#endif
