"""`-m gpu`: the reference's known-answer tests against the real CUDA build through the C ABI."""
import pytest

from tests import kat

KATS = [getattr(kat, n) for n in sorted(dir(kat)) if n.startswith("kat_")]
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fn", KATS, ids=[f.__name__ for f in KATS])
def test_kat_gpu(product, fn):
    fn(product)


def test_extension_is_the_cuda_build(product):
    assert b"sm_100a" in product.api.version()
    assert product.api.device_count() >= 1
    before = product.api.kernel_launch_count()
    kat.kat_detects_sine_frequency_peak(product)
    assert product.api.kernel_launch_count() > before


def test_unsupported_size_is_an_error_not_a_fallback(product):
    """Non power-of-two sizes are legal in the reference (rustfft is mixed radix); this build has no kernel
    for them and no CPU fallback: explicit OMB_ERR_UNSUPPORTED."""
    import numpy as np

    from openmeters_b200.processors import AudioBlock, OmbError, SpectrogramConfig

    p = product.Spectrogram(SpectrogramConfig(fft_size=1000, hop_size=250, use_reassignment=False))
    with pytest.raises(OmbError, match="power-of-two"):
        p.process_block(AudioBlock(np.zeros(2000, np.float32), 1, 48000.0))
