"""The drop-in check at the reference's own call site (oracle/adapter_check.cpp): reference SDPAlign candidates refined by
the reference's (Affine)GuidedAlign + ComputeAlignmentStats and by include/blasr_gpu_adapter.hpp into the reference's real
T_AlignmentCandidate.  The binary is prebuilt by oracle/Makefile (it needs /root/reference at build time only)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adapter_check")

needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/adapter_check not built (needs /root/reference)")


@needs_bin
@pytest.mark.gpu
def test_adapter_matches_reference_call_site():
    r = subprocess.run([BIN, "64", "8000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical to the reference call site") == 2, r.stdout
    # SAM CIGAR ('=' / 'X' / 'I' / 'D') and the m5 strings, printed by the reference's own printers from both candidates
    assert r.stdout.count("printed by the reference's printers: identical") == 2, r.stdout


@needs_bin
@pytest.mark.gpu
def test_dense_adapter_matches_reference_call_sites():
    """KBandAlign (Global, Fit), SWAlign (Global) and AffineKBandAlign (Global) through blasr_gpu::DenseBatch against the
    reference's templates called with blasr's own argument patterns."""
    r = subprocess.run([BIN, "32", "3000", "dense"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("through blasr_gpu::DenseBatch: identical to the reference call site") == 4, r.stdout


@needs_bin
@pytest.mark.gpu
def test_sdp_adapter_matches_reference_call_site():
    """SDPAlign through blasr_gpu::SdpBatch into the reference's real T_AlignmentCandidate, against the reference's own
    SDPAlign called with blasr's argument lists (Blasr.cpp:1716-1722 Local, :1080-1090 Global)."""
    r = subprocess.run([BIN, "48", "6000", "sdp"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "through blasr_gpu::SdpBatch: identical to the reference call site" in r.stdout, r.stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "anchor_check")), reason="oracle/_ref/anchor_check not built")
@pytest.mark.gpu
def test_anchor_adapter_matches_reference_call_site():
    """MapReadToGenome through blasr_gpu::AnchorBatch, loaded from the reference's real DNASuffixArray / DNASequence and fed
    its SMRTSequence reads (forward and MakeRC, whole reads and subreads), against the reference's own calls
    (Blasr.cpp:2282-2296) into vector<ChainedMatchPos>."""
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "anchor_check"), "24", "400000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "through blasr_gpu::AnchorBatch: identical to the reference call site" in r.stdout, r.stdout


@needs_bin
def test_adapter_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([BIN, "2", "400"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
