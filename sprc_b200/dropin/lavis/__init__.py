"""Drop-in `lavis` package: put `sprc_b200/dropin` on PYTHONPATH and the reference's unchanged
`src/blip_validate.py`, `src/cirr_test_submission.py` and `src/validate_blip.py` import
`lavis.models.load_model_and_preprocess` from here instead of the vendored LAVIS."""
