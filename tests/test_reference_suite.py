"""The reference's OWN model tests (tests/test_two_tower_model.py, 19 tests: shapes, unit norms, gradient flow, losses,
save/load, the factory, embedding quality after a short training), run UNMODIFIED against b200rec on the GPU.  The file
is staged by baseline/stage_reference.py (git-ignored copy; /root/reference does not exist on the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST = os.path.join(ROOT, "baseline", "_ref", "tests", "test_two_tower_model.py")


@pytest.mark.gpu
def test_reference_model_tests_pass_unmodified_on_b200rec():
    if not os.path.isfile(REF_TEST):
        pytest.skip("reference tests not staged (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TEST, "-q", "-p", "ref_suite_plugin", "-o", "addopts=",
                        "-p", "no:cacheprovider", "--rootdir", os.path.dirname(os.path.dirname(REF_TEST)), "-c", os.devnull],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "19 passed" in r.stdout, tail


def test_staged_reference_is_not_tracked():
    """baseline/_ref/ must stay out of the history (it holds reference sources)."""
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "baseline/_ref/" in gi
    r = subprocess.run(["git", "ls-files", "baseline/_ref"], cwd=ROOT, capture_output=True, text=True)
    assert r.stdout.strip() == ""
