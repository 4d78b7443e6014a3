"""Drop-in mirror of the reference's `contrast` package for the pixel-pretext hot path.

Same import paths, function names, argument meaning and error behaviour as the reference
(contrast.models.PixPro, contrast.util, contrast.flow.upflow8); the arithmetic runs in the
sm_100a kernels of libpixpro_b200.so.  Out-of-scope subsystems of the reference (data
pipeline, RAFT estimator, LARS, logging, linear eval) are not mirrored — SURVEY.md §2.
"""
