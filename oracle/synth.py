"""TEST INFRASTRUCTURE — the synthetic checkpoint / input generators live in sprc_b200/synth.py (bench.py's own arm
needs them and may not import from oracle/); this module re-exports them for the oracle, the golden-vector generator
and the tests."""
from sprc_b200.synth import *  # noqa: F401,F403
from sprc_b200.synth import VIT_DIMS, make_dyadic, make_gallery_features, make_images, make_state_dict  # noqa: F401
from sprc_b200.synth import make_token_ids, state_dict_spec  # noqa: F401
