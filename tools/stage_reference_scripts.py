"""Copies the reference's five retrieval driver scripts BYTE-FOR-BYTE into the git-ignored staging
directory baseline/_ref/src/ (SURVEY.md §8b "How unchanged is realised"): data_utils.base_path is the
parent of the directory holding data_utils.py and /root/reference is read-only, so the unchanged
scripts must run from a writable copy next to a synthetic dataset tree.  Nothing is modified and
nothing lands in git history; `lavis/` is NOT staged — PYTHONPATH supplies sprc_b200/dropin instead.
Run in the build container (build() calls it when /root/reference exists)."""
import hashlib
import os
import shutil
import sys

FILES = ["blip_validate.py", "cirr_test_submission.py", "validate_blip.py", "utils.py", "data_utils.py"]


def stage(reference_src="/root/reference/src", root=None):
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "baseline", "_ref", "src")
    if not os.path.isdir(reference_src):
        return None
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(reference_src, f), os.path.join(dst, f))
        a = hashlib.sha256(open(os.path.join(reference_src, f), "rb").read()).hexdigest()
        b = hashlib.sha256(open(os.path.join(dst, f), "rb").read()).hexdigest()
        assert a == b
    return dst


if __name__ == "__main__":
    print(stage(*(sys.argv[1:2])))
