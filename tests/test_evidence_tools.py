"""The tools that turn ncu output into what profiles/ holds keep working on the committed captures, and the in-tree
library really is sm_100a code with the instructions DESIGN.md talks about (cuobjdump works without a GPU)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.check_output([sys.executable] + list(args), cwd=ROOT, text=True, stderr=subprocess.STDOUT)


def test_launch_list_of_the_fused_frame_parses():
    out = _run("tools/launch_frame.py", "profiles/r2final/launches_frame_warm_kitti_final.csv")
    for kernel in ("repitch_kernel", "fast_nms_kernel", "compact_frame_kernel", "describe_tile_kernel", "track_search_kernel",
                   "track_resolve_kernel", "track_emit_kernel", "converge_cluster_kernel", "match_kernel",
                   "select_strips_kernel", "frame_assemble_kernel", "total"):
        assert kernel in out, kernel
    total = float([l for l in out.splitlines() if l.startswith("total")][0].split()[1])
    assert 80 < total < 250          # microseconds per tracked KITTI frame, serialised


def test_launch_summary_of_the_bench_step_parses():
    out = _run("tools/ncu_summary.py", "launches", "profiles/r2final/launches_bench_pairs512.csv")
    assert "kernels of the bench step only" in out
    share = {l.split()[0]: float(l.split()[-1]) for l in out.split("kernels of the bench step only:")[1].splitlines() if l.strip()}
    assert share["fast_nms_kernel"] > share["blur_kernel"] > share["match_kernel"]
    assert abs(sum(share.values()) - 1.0) < 0.01


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="no cuobjdump")
def test_the_library_is_sm_100a_code_with_tma_and_packed_math():
    out = _run("tools/sass_summary.py")
    assert "arch sm_100a" in out
    blocks = {b.split()[0]: b for b in out.split("\n\n") if b.strip() and not b.startswith("#")}
    assert "UTMALDG" in blocks["fast_nms_kernel"] and "VIMNMX3.S16" in blocks["fast_nms_kernel"]
    assert "FFMA2" in blocks["blur_kernel"] and "UTMALDG" in blocks["blur_kernel"]
    assert "UTMALDG" in blocks["describe_tile_kernel"]
    cluster = [b for k, b in blocks.items() if k.startswith("converge_cluster_kernel")]
    assert cluster and all("DFMA" in b or "DADD" in b for b in cluster)
    assert any(k.startswith("compact_frame_kernel") for k in blocks) and any(k.startswith("track_emit_kernel") for k in blocks)
