"""Stage the UNMODIFIED reference files the GPU box needs under baseline/_ref/ (git-ignored, shipped by gpurun).

/root/reference does not exist on the GPU box, but three things there must run the reference's own code:
  * bench.py's `cpu_baseline` / `--impl reference` training legs: the imported reference TwoTowerModel + the step body of
    its TwoTowerTrainer on torch-CPU;
  * bench.py's config-1 epoch: the reference's MovieLens loader (src/data/movielens.py) over ml-1m/{users,movies}.dat
    and a synthesised ratings.dat;
  * tests/test_reference_suite.py: the reference's tests/test_two_tower_model.py, unmodified, against b200rec on CUDA.
Nothing here is product source and nothing is copied into the tracked tree: baseline/_ref/ is listed in .gitignore.
`__graft_entry__.build()` calls stage() whenever /root/reference is present (i.e. in the build container)."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
WANT = ["src/constants.py", "src/models", "src/training", "src/data", "src/evaluation", "scripts/evaluate_model.py",
        "tests/test_two_tower_model.py", "tests/conftest.py", "ml-1m/users.dat", "ml-1m/movies.dat", "pytest.ini"]


def stage(reference: str = "/root/reference") -> bool:
    if not os.path.isdir(reference):
        return os.path.isdir(os.path.join(DST, "src", "models"))
    for rel in WANT:
        src, dst = os.path.join(reference, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(src, dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "models", "two_tower.py"))


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")
