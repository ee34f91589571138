"""Copies the reference's five retrieval driver scripts and the 14 model-side files its
`blip2_cir_align_prompt` / `blip2_cir_rerank` classes import (MODEL_FILES; what `oracle/ref_loader.py` loads, found by
listing sys.modules after the import) BYTE-FOR-BYTE into the git-ignored staging
directory baseline/_ref/src/ (SURVEY.md §8b "How unchanged is realised"): data_utils.base_path is the
parent of the directory holding data_utils.py and /root/reference is read-only, so the unchanged
scripts must run from a writable copy next to a synthetic dataset tree.  Nothing is modified and
nothing lands in git history.  The drop-in tests put sprc_b200/dropin FIRST on PYTHONPATH, so `import lavis` there still
resolves to our package; the staged `lavis/` tree has no __init__.py files and is only reachable through
oracle/ref_loader.py (SPRC_REFERENCE_SRC), which bench.py's `--impl reference` arm and its `eager_gpu` row use to run the
UNMODIFIED reference model on the GPU box, where /root/reference does not exist.
Run in the build container (build() calls it when /root/reference exists)."""
import hashlib
import os
import shutil
import sys

FILES = ["blip_validate.py", "cirr_test_submission.py", "validate_blip.py", "utils.py", "data_utils.py"]
MODEL_FILES = [
    "lavis/common/dist_utils.py", "lavis/common/logger.py", "lavis/common/registry.py", "lavis/common/utils.py",
    "lavis/models/base_model.py", "lavis/models/clip_vit.py", "lavis/models/eva_vit.py",
    "lavis/models/blip2_models/Qformer.py", "lavis/models/blip2_models/blip2.py",
    "lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py",
    "lavis/models/blip2_models/blip2_qformer_cir_rerank.py",
    "lavis/models/blip_models/blip_outputs.py",
    "lavis/processors/base_processor.py", "lavis/processors/blip_processors.py",
]


def stage(reference_src="/root/reference/src", root=None):
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "baseline", "_ref", "src")
    if not os.path.isdir(reference_src):
        return None
    os.makedirs(dst, exist_ok=True)
    for f in FILES + MODEL_FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(reference_src, f), os.path.join(dst, f))
        a = hashlib.sha256(open(os.path.join(reference_src, f), "rb").read()).hexdigest()
        b = hashlib.sha256(open(os.path.join(dst, f), "rb").read()).hexdigest()
        assert a == b
    return dst


if __name__ == "__main__":
    print(stage(*(sys.argv[1:2])))
