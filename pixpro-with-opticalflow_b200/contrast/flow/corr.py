"""contrast.flow.corr — drop-in `CorrBlock` of the reference's RAFT estimator (contrast/flow/corr.py:12-60) on this
package's kernels (csrc/pp_corr.cu): the all-pairs correlation volume is one tcgen05 3xTF32 batched contraction with
the 1/sqrt(dim) division in its epilogue, every pyramid level one pooling launch, and a lookup (`__call__`) ONE launch
for all levels that writes the [B, L*(2r+1)^2, h, w] result directly — the reference's per-level meshgrid, broadcast
add, `grid_sample`, `view`, `cat`, `permute`, `contiguous` chain (7 torch ops and a [B*h*w, 2r+1, 2r+1, 2] coordinate
tensor per level and iteration) is gone.  The reference's CUDA twin (`alt_cuda_corr`, corr.py:63-91) is an extension
it does not ship; `AlternateCorrBlock` is therefore not provided here either.
"""
from pixpro_b200 import ops as _ops


class CorrBlock:
    """Same constructor, attributes and call convention as the reference class.

    corr_pyramid[l] has the reference's shape [B*h*w, 1, h >> l, w >> l]; `__call__(coords)` takes [B,2,h,w] pixel
    coordinates (x, y) of level 0 and returns float32 [B, num_levels*(2*radius+1)**2, h, w]."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        batch, dim, ht, wd = fmap1.shape
        corr = CorrBlock.corr(fmap1, fmap2).reshape(batch * ht * wd, 1, ht, wd)  # corr.py:21-22 (a view)
        self.corr_pyramid = [corr]
        for _ in range(num_levels - 1):
            corr = _ops.corr_pool(corr)                                             # corr.py:26-28
            self.corr_pyramid.append(corr)

    def __call__(self, coords):
        return _ops.corr_lookup(self.corr_pyramid, coords, self.radius)             # corr.py:30-50

    @staticmethod
    def corr(fmap1, fmap2):
        """corr.py:52-60: [B,dim,h,w] x2 -> [B,h,w,1,h,w] = <fmap1[:, :, i], fmap2[:, :, j]> / sqrt(dim)."""
        batch, dim, ht, wd = fmap1.shape
        return _ops.corr_volume(fmap1.float(), fmap2.float()).view(batch, ht, wd, 1, ht, wd)


__all__ = ['CorrBlock']
