"""Multi-GPU pool behind the C ABI (xdtts_pool_*): the result of a pooled call does not depend on
the number of devices -- bit for bit against single-device calls."""
import numpy as np
import pytest

from oracle import gl_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gl(built):
    from xdtts_b200 import griffin_lim

    return griffin_lim


def _basis():
    return o.create_mel_filter_bank(22050.0, 1024, 80, 0.0, 8000.0)


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2])
def test_pool_equals_single_device_bitwise(gl, n_dev):
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    ts = [40, 9, 123, 5, 64, 17, 33]
    mels = [o.synth_mel(300 + i, 80, t) for i, t in enumerate(ts)]
    phs = [o.phase_turns(9, i, 513, t) for i, t in enumerate(ts)]
    # a fixed run length makes the result independent of how the batch is split (see include/xdtts_b200.h)
    one = gl.GriffinLim.new(_basis(), 768, 1.7, 6, 0.99, run_frames=8, seed=21, fixed_seed=True)
    pool = gl.GriffinLimPool.new(_basis(), 768, 1.7, 6, 0.99, devices=list(range(n_dev)), run_frames=8, seed=21, fixed_seed=True)
    assert pool.n_devices() == n_dev
    want_explicit = one.infer_batch(mels, phs)
    want_seeded = one.infer_batch(mels)
    for got, want in ((pool.infer_batch(mels, phs), want_explicit), (pool.infer_batch(mels), want_seeded)):
        assert len(got) == len(ts)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    # the seeded stream of utterance i is keyed by its index in the caller's batch, whichever device vocodes it
    for i in (0, 3):
        (alone,) = one.infer_batch([mels[i]], [o.phase_turns(21, i, 513, ts[i])])
        assert np.array_equal(alone, want_seeded[i])
    # assignment: longest first, least-loaded device, ties to the first device
    dev = pool.assignment(ts)
    assert len(dev) == len(ts) and set(dev) <= set(range(n_dev))
    if n_dev == 2:
        load = [sum(t for t, d in zip(ts, dev) if d == k) for k in range(2)]
        assert abs(load[0] - load[1]) <= max(ts)
    # magnitudes in, and the error path (message carried over from the worker thread)
    mags = [o.lift_pinv_clamp(m, _basis(), 1.7, dtype=np.float32) for m in mels[:3]]
    for a, b in zip(pool.from_magnitude_batch(mags, phs[:3]), one.from_magnitude_batch(mags, phs[:3])):
        assert np.array_equal(a, b)
    from xdtts_b200._ffi import ERR_SHAPE, XdttsError

    with pytest.raises(XdttsError) as e:
        pool.infer_batch([o.synth_mel(1, 80, 1)])                 # one frame = no samples
    assert e.value.code == ERR_SHAPE
    pool.close()
    one.close()


def test_pool_new_phase_field_per_call(gl):
    pool = gl.GriffinLimPool.new(_basis(), 768, 1.7, 3, 0.99, devices=[0], seed=4)
    mel = [o.synth_mel(2, 80, 20)]
    a, b = pool.infer_batch(mel)[0], pool.infer_batch(mel)[0]
    assert not np.array_equal(a, b)                            # like the reference: new random phases per call
    pool.close()
