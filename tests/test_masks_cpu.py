"""The product's mask generators (viai_b200/utils/masks.py) against the oracle's restatements: bit identical."""
import pytest
import torch

from oracle import viai_oracle as O


@pytest.mark.parametrize("shape,seed", [((1, 1, 80, 64), 0), ((3, 1, 128, 128), 128), ((2, 1, 256, 256), 7), ((1, 1, 512, 512), 512)])
def test_freeform_mask_is_bit_identical_to_the_oracle(shape, seed):
    from viai_b200.utils.masks import freeform_mask
    got, want = freeform_mask(shape, seed), O.freeform_mask(shape, seed)
    assert got.dtype == torch.float32 and torch.equal(got, want)
    assert set(got.unique().tolist()) <= {0.0, 1.0} and 0.0 < float((got == 0).float().mean()) < 0.9


def test_time_band_mask_is_bit_identical_to_the_oracle():
    from viai_b200.utils.masks import time_band_mask
    for shape, t0, bl in (((2, 1, 80, 64), 16, 32), ((1, 1, 256, 256), 64, 128), ((1, 1, 80, 64), 60, 32), ((1, 1, 80, 64), 0, 0)):
        assert torch.equal(time_band_mask(shape, t0, bl), O.time_band_mask(shape, t0, bl))
