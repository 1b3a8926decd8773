import numpy as np
import pytest

from nkb_classification_b200 import transforms as T
from oracle import preprocess as opre


def test_compile_resize():
    p = T.compile_pipeline(T.Compose([T.Resize(224, 200), T.Normalize(), T.ToTensorV2()]))
    assert (p.mode, p.out_h, p.out_w) == (T.MODE_STRETCH, 224, 200)
    m, d = opre.normalize_constants((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert np.array_equal(np.array(p.mean255, np.float32), m) and np.array_equal(np.array(p.denom, np.float32), d)


def test_compile_letterbox_like_reference_configs():
    # configs/singletask_config.py:203-219
    p = T.compile_pipeline(T.Compose([
        T.LongestMaxSize(128, always_apply=True),
        T.PadIfNeeded(128, 128, always_apply=True, border_mode=T.BORDER_CONSTANT, value=0),
        T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
        T.ToTensorV2(),
    ]))
    assert (p.mode, p.out_h, p.out_w, p.max_size, p.pad_value) == (T.MODE_LETTERBOX, 128, 128, 128, (0, 0, 0))


def test_duck_typed_albumentations_objects():
    class Resize:  # stands in for albumentations.Resize (same class name + public attrs)
        height, width, interpolation = 64, 32, 1

    class Normalize:
        mean, std, max_pixel_value = (0.5, 0.5, 0.5), (0.25, 0.25, 0.25), 255.0

    class ToTensorV2:
        pass

    class Compose:
        transforms = [Resize(), Normalize(), ToTensorV2()]

    p = T.compile_pipeline(Compose())
    assert (p.out_h, p.out_w) == (64, 32)


@pytest.mark.parametrize("ops", [
    [T.Resize(8, 8), T.ToTensorV2()],                                            # no Normalize
    [T.Resize(8, 8), T.Normalize()],                                             # no ToTensorV2
    [T.LongestMaxSize(8), T.Normalize(), T.ToTensorV2()],                        # variable output size
    [T.LongestMaxSize(8), T.PadIfNeeded(8, 8), T.Normalize(), T.ToTensorV2()],   # default BORDER_REFLECT_101
    [T.LongestMaxSize(16), T.PadIfNeeded(8, 8, border_mode=0), T.Normalize(), T.ToTensorV2()],
    [T.Resize(8, 8, interpolation=2), T.Normalize(), T.ToTensorV2()],            # cubic
    [type("HorizontalFlip", (), {})(), T.Resize(8, 8), T.Normalize(), T.ToTensorV2()],  # random train-time op
])
def test_unsupported_pipelines_raise(ops):
    with pytest.raises(NotImplementedError):
        T.compile_pipeline(ops)
