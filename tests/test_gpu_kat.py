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


def test_unsupported_update_keeps_the_previous_config(product):
    """update_config to a size without a kernel must not leave a prepared handle that fails every later block:
    the update is refused (OMB_ERR_UNSUPPORTED) and the handle keeps working with its previous configuration."""
    import numpy as np

    from openmeters_b200.processors import AudioBlock, OmbError, SpectrogramConfig, SpectrumConfig

    x = np.sin(np.arange(4096, dtype=np.float32) * np.float32(0.05))
    p = product.Spectrogram(SpectrogramConfig(fft_size=1024, hop_size=256, use_reassignment=False, history_length=16))
    assert p.process_block(AudioBlock(x, 1, 48000.0)) is not None
    c = p.config()
    c.fft_size = 1000
    with pytest.raises(OmbError, match="previous config kept"):
        p.update_config(c)
    assert p.config().fft_size == 1024
    assert p.process_block(AudioBlock(x, 1, 48000.0)) is not None
    s = product.Spectrum(SpectrumConfig(fft_size=1024, hop_size=256))
    assert s.process_block(AudioBlock(x, 1, 48000.0)) is not None
    sc = s.config()
    sc.fft_size = 1500
    with pytest.raises(OmbError, match="previous config kept"):
        s.update_config(sc)
    assert s.config().fft_size == 1024 and s.process_block(AudioBlock(x, 1, 48000.0)) is not None
