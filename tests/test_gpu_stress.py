"""The decoupled kernel's protocol under forced re-orderings: the jittered build (random sleeps at every hand-off, one
draw per lane) must finish and equal the first-generation kernel byte for byte (tools/repro_race.py).

Two faults of this kernel were timing dependent and invisible to every ordinary parity test -- a miscompute in round 1, a
wait for a barrier phase that never comes in round 2 -- so the stress build is part of the GPU suite."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
VARIANT = ROOT / "build" / "variants" / "libhvx_jitter2.so"


@pytest.mark.parametrize("shape", [
    ["--edge", "32", "--chunks", "140"],                      # split walk (fewer chunks than resident CTAs), partially dirty chunks
    ["--edge", "32", "--chunks", "900"],                      # several chunks per CTA, three CTAs per SM
    ["--edge", "64", "--chunks", "60"],                       # split walk at edge 64
    ["--edge", "64", "--chunks", "300", "--mixed"],           # surface and empty chunks mixed
    ["--edge", "64", "--chunks", "256", "--sparse-dirty"],    # an edit frame: most chunks skipped, a few partially dirty
])
def test_jittered_build_finishes_and_matches(shape):
    if not VARIANT.exists():
        pytest.skip("build/variants/libhvx_jitter2.so is not built (make -C helio_b200/csrc ../../build/variants/libhvx_jitter2.so)")
    cmd = [sys.executable, str(ROOT / "tools" / "repro_race.py"), "--lib", str(VARIANT.relative_to(ROOT)), "--iters", "6", "--full-every", "3", *shape]
    try:
        proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired:
        pytest.fail("the jittered launch did not finish (tools/wait_trace.py shows what it waits for)")
    last = (proc.stdout.strip().splitlines() or [""])[-1]
    assert proc.returncode == 0 and "bad_runs=0 invariant_violations=0" in last, proc.stdout[-2000:] + proc.stderr[-2000:]
