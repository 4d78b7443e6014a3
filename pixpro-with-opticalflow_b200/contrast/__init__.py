"""Drop-in mirror of the reference's `contrast` package for the pixel-pretext hot path.

Same import paths, function names, argument meaning and error behaviour as the reference
(contrast.models.PixPro, contrast.util, contrast.flow.upflow8); the arithmetic runs in the
sm_100a kernels of libpixpro_b200.so.  Out-of-scope subsystems of the reference (data
pipeline, RAFT estimator, lr scheduler, logging, linear eval) are not mirrored — SURVEY.md §2.

Mixing with the reference tree: put this repo's package directory BEFORE the reference on sys.path.
`contrast` then resolves here, and every sub-module this package does not provide (`contrast.data`,
`contrast.option`, `contrast.lr_scheduler`, `contrast.logger`, ...) is found in the reference's own
`contrast/` directory, which is appended to this package's search path below.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
