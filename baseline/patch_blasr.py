#!/usr/bin/env python3
"""baseline/patch_blasr.py -- edits a COPY of the reference's alignment/Blasr.cpp at build time (baseline/Makefile).

    patch_blasr.py dump <Blasr.cpp> <out.cpp>   job dump at the refinement call sites (baseline/job_dump.hpp)
    patch_blasr.py gpu  <Blasr.cpp> <out.cpp>   RefineAlignments routed through include/blasr_gpu_adapter.hpp
                                                (baseline/gpu_refine.hpp = the patch of INTEGRATION.md section 2)

Every edit is an insertion anchored on text that must occur exactly once; the script fails loudly when the
reference differs from what it expects.  No reference text is stored here, and the patched copy only exists
under baseline/_ref/build/ (git-ignored).
"""
import sys


def insert_before(src: str, anchor: str, text: str, nth_line_back: int = 0) -> str:
    n = src.count(anchor)
    if n != 1:
        raise SystemExit(f"patch_blasr: anchor {anchor!r} occurs {n} times (expected 1)")
    pos = src.index(anchor)
    for _ in range(nth_line_back + 1):            # start of the line that holds the anchor (or lines above it)
        pos = src.rfind("\n", 0, pos)
    pos += 1
    return src[:pos] + text + src[pos:]


def insert_after_line(src: str, anchor: str, text: str) -> str:
    n = src.count(anchor)
    if n != 1:
        raise SystemExit(f"patch_blasr: anchor {anchor!r} occurs {n} times (expected 1)")
    pos = src.index("\n", src.index(anchor)) + 1
    return src[:pos] + text + src[pos:]


def dump(src: str) -> str:
    # the include goes right before the first function that uses it
    src = insert_before(src, "void RefineAlignment(T_Sequence &query,", '#include "job_dump.hpp"\n', nth_line_back=1)
    # RefineAlignment, Blasr.cpp:862: just before `if (params.affineAlign) { AffineGuidedAlign(qSeq, tSeq, ...`
    src = insert_before(src, "AffineGuidedAlign(qSeq, tSeq, alignmentCandidate,",
                        "      bgpu_dump::Guided(qSeq, tSeq, alignmentCandidate, distScoreFn,\n"
                        "                        params.affineAlign ? params.bandSize : params.guidedAlignBandSize, params.affineAlign);\n",
                        nth_line_back=1)
    # AlignSubstring, Blasr.cpp:1067: just before `alignScore = AffineKBandAlign(qSubSeq, tSubSeq, ...`
    src = insert_before(src, "alignScore = AffineKBandAlign(qSubSeq, tSubSeq, distScoreFn.scoreMatrix,",
                        "\t\tbgpu_dump::AffineKBand(qSubSeq, tSubSeq, distScoreFn.scoreMatrix, params.indel+2, params.indel-3,\n"
                        "\t\t                       params.indel+2, params.indel-1, params.indel, params.bandSize);\n")
    return src


def replace_once(src: str, old: str, new: str) -> str:
    n = src.count(old)
    if n != 1:
        raise SystemExit(f"patch_blasr: text {old!r} occurs {n} times (expected 1)")
    return src.replace(old, new)


def gpu(src: str) -> str:
    # gpu_refine.hpp needs T_AlignmentCandidate, MappingParameters, MappingBuffers, SortAlignmentPointersByScore: all are
    # defined above RefineAlignments (Blasr.cpp:2163)
    src = insert_before(src, "void RefineAlignments(vector<T_Sequence*> &bothQueryStrands,", '#include "gpu_refine.hpp"\n',
                        nth_line_back=1)
    # first statement of RefineAlignments: the batched GPU path takes the whole candidate list of this read
    src = insert_after_line(src, "vector<T_AlignmentCandidate*> &alignmentPtrs, MappingParameters &params, MappingBuffers &mappingBuffers) {",
                            "  if (BgpuRefineAlignments(bothQueryStrands, genome, alignmentPtrs, params, mappingBuffers)) return;\n")
    # thread driver, Blasr.cpp:4838-4841: the -nproc MapReads instances run as fibers on one pthread per core
    src = replace_once(src, "pthread_create(&threads[procIndex], &threadAttr[procIndex], (void* (*)(void*))MapReads, &mapdb[procIndex]);",
                       "BgpuSpawn(&threads[procIndex], &threadAttr[procIndex], (void* (*)(void*))MapReads, &mapdb[procIndex], procIndex, params.nProc);")
    src = replace_once(src, "pthread_join(threads[procIndex], NULL);", "BgpuJoin(threads[procIndex], procIndex);")
    src = replace_once(src, "pthread_exit(NULL);", "BgpuFiberExit();")                      # end of MapReads, Blasr.cpp:3915
    # anchoring, Blasr.cpp:2282-2296: both strands in one device call (the second call site takes the list the first one filled)
    src = replace_once(src, "MapReadToGenome(genome, sarray, read, ",
                       "BgpuMapReadToGenome(&readRC, &mappingBuffers.rcMatchPosList, params.forwardOnly, genome, sarray, read, ")
    src = replace_once(src, "MapReadToGenome(genome, sarray, readRC, params.lookupTableLength, mappingBuffers.rcMatchPosList,",
                       "BgpuMapReadToGenomeRC(genome, sarray, readRC, params.lookupTableLength, mappingBuffers.rcMatchPosList,")
    return src


def main():
    mode, inp, out = sys.argv[1:4]
    src = open(inp, encoding="latin-1").read()
    src = {"dump": dump, "gpu": gpu}[mode](src)
    open(out, "w", encoding="latin-1").write(src)


if __name__ == "__main__":
    main()
