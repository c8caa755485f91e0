"""Multi-GPU paths on a box with at least two B200s (pytest -m gpu; skipped on a one-GPU box): the device-group API
(host-sharded and device-resident with the fused peak all-gather + row gather) and the two-PROCESS form of the gather
(pdsp_spectrum_dev_gather over CUDA-IPC peer buffers), every gathered byte checked against an NCCL all_gather."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")


def multitone(rng, batch, n, dtype=np.float64):
    t = np.arange(n)
    k = rng.integers(8, n // 2 - 8, size=(batch, 3)) + rng.uniform(-0.25, 0.25, size=(batch, 3))
    a = np.concatenate([np.ones((batch, 1)), rng.uniform(0.1, 0.5, size=(batch, 2))], axis=1)
    ph = rng.uniform(0, 2 * np.pi, size=(batch, 3))
    x = np.zeros((batch, n))
    for j in range(3):
        x += a[:, j, None] * np.sin(2 * np.pi * k[:, j, None] * t[None, :] / n + ph[:, j, None])
    return x.astype(dtype)


@needs2
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_group_host_sharded_equals_one_device(prec):
    import oracle
    from pragma_dsp_b200 import DeviceGroup, spectrum_batch
    rng = np.random.default_rng(5)
    n, batch = 1024, 30001  # uneven blocks, several staging chunks per device
    x = multitone(rng, batch, n, np.float64 if prec == "f64" else np.float32)
    one = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision=prec, outputs=("amplitude", "peak"))
    with DeviceGroup(list(range(min(_ngpu(), 8)))) as grp:
        got = grp.spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann", precision=prec, outputs=("amplitude", "peak"))
    assert np.array_equal(got["amplitude"], one["amplitude"]) and (got["peaks"] == one["peaks"]).all()
    ref = oracle.spectrum_batch(x[:512], fftSize=n, sampleRate=48000.0, window="hann")
    assert (got["peaks"]["index"][:512] == ref["peaks"]["index"]).all()


@needs2
def test_group_device_resident_gather():
    import torch

    from pragma_dsp_b200 import DeviceGroup, spectrum_batch
    from pragma_dsp_b200._lib import F64, PEAK_F64, SIDES, WINDOWS, SpectrumDesc
    from pragma_dsp_b200.group import block_of
    g = min(_ngpu(), 8)
    rng = np.random.default_rng(6)
    n, batch = 1024, 4099
    bins = n // 2 + 1
    x = multitone(rng, batch, n)
    ref = spectrum_batch(x, sampleRate=48000.0, fftSize=n, window="hann")
    blocks = [block_of(batch, g, i) for i in range(g)]
    d_x = [torch.from_numpy(x[f0:f0 + nf]).to(f"cuda:{i}") for i, (f0, nf) in enumerate(blocks)]
    d_amp = [torch.empty((nf, bins), dtype=torch.float64, device=f"cuda:{i}") for i, (_, nf) in enumerate(blocks)]
    d_ph = [torch.empty((nf, bins), dtype=torch.float64, device=f"cuda:{i}") for i, (_, nf) in enumerate(blocks)]
    d_pk = [torch.zeros((batch, 32), dtype=torch.uint8, device=f"cuda:{i}") for i in range(g)]
    root = g - 1
    amp_all = torch.empty((batch, bins), dtype=torch.float64, device=f"cuda:{root}")
    ph_all = torch.empty((batch, bins), dtype=torch.float64, device=f"cuda:{root}")
    for i in range(g):
        torch.cuda.synchronize(i)
    desc = SpectrumDesc(sample_dtype=F64, frame_len=n, hop=n, batch=batch, window=WINDOWS["hann"], sides=SIDES["one"],
                        sample_rate=48000.0, raw_magnitude=0, fft_shift=0)
    ptr = lambda ts: [t.data_ptr() for t in ts]  # noqa: E731
    with DeviceGroup(list(range(g))) as grp:
        for _ in range(2):
            grp.spectrum_dev(n, F64, desc, ptr(d_x), ptr(d_amp), ptr(d_ph), ptr(d_pk), gather_root=root,
                             d_amplitude_all=amp_all.data_ptr(), d_phase_all=ph_all.data_ptr())
            grp.sync()
        for i in range(g):
            rec = d_pk[i].cpu().numpy().view(PEAK_F64).reshape(-1)
            assert (rec == ref["peaks"]).all(), f"device {i} does not hold every frame's record"
        assert np.array_equal(amp_all.cpu().numpy(), ref["amplitude"]) and np.array_equal(ph_all.cpu().numpy(), ref["phase"])


def _ipc_worker(rank, world, port, frames, q):
    try:
        import torch
        import torch.distributed as dist

        from pragma_dsp_b200 import _lib
        from pragma_dsp_b200._lib import F32, PEAK_F32, SIDES, WINDOWS, SpectrumDesc, check, lib
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        L = lib()
        ctx = _lib.Context(rank)
        n, pkb = 1024, 16
        rng = np.random.default_rng(100 + rank)
        x = torch.from_numpy(multitone(rng, frames, n, np.float32)).to(dev)
        plan = ctx.plan(n, F32)
        gbuf = C.c_void_p()
        check(L.pdsp_dev_alloc(ctx.h, world * frames * pkb, C.byref(gbuf)))
        hb = C.create_string_buffer(64)
        check(L.pdsp_ipc_export(ctx.h, gbuf, hb))
        handles = [None] * world
        dist.all_gather_object(handles, hb.raw)
        peers = (C.c_void_p * 8)()
        for r in range(world):
            if r == rank:
                peers[r] = gbuf.value
            else:
                pp = C.c_void_p()
                check(L.pdsp_ipc_open(ctx.h, handles[r], C.byref(pp)))
                peers[r] = pp.value
        dist.barrier()
        local = torch.zeros((frames, pkb), dtype=torch.uint8, device=dev)
        desc = SpectrumDesc(sample_dtype=F32, frame_len=n, hop=n, batch=frames, window=WINDOWS["hann"], sides=SIDES["one"],
                            sample_rate=48000.0, raw_magnitude=0, fft_shift=0)
        st = torch.cuda.Stream(device=dev)
        st.wait_stream(torch.cuda.current_stream())  # ordered behind the generators / fills on torch's stream
        check(L.pdsp_spectrum_dev_gather(plan, C.byref(desc), C.c_void_p(x.data_ptr()), None, None, C.c_void_p(local.data_ptr()),
                                         peers, world, rank * frames, C.c_void_p(st.cuda_stream)))
        st.synchronize()
        dist.barrier()  # every rank's kernel has finished its peer stores
        torch.cuda.synchronize()
        ref = torch.empty((world * frames, pkb), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(ref, local)
        host = np.zeros(world * frames * pkb, dtype=np.uint8)
        check(L.pdsp_memcpy_d2h(ctx.h, C.c_void_p(host.ctypes.data), gbuf, host.nbytes, C.c_void_p(st.cuda_stream)))
        st.synchronize()
        same = bool((host.reshape(world * frames, pkb) == ref.cpu().numpy()).all())
        idx_ok = bool((host.view(PEAK_F32)["index"] >= 8).all())
        dist.barrier()
        for r in range(world):
            if r != rank:
                check(L.pdsp_ipc_close(ctx.h, C.c_void_p(peers[r])))
        dist.barrier()
        check(L.pdsp_dev_free(ctx.h, gbuf))
        dist.destroy_process_group()
        q.put((rank, same and idx_ok, ""))
    except Exception as e:  # pragma: no cover
        q.put((rank, False, repr(e)))


@needs2
def test_two_process_ipc_gather_is_byte_exact():
    """pdsp_spectrum_dev_gather + pdsp_ipc_*: two processes, one GPU each; EVERY segment of EVERY rank's gathered buffer
    equals what an NCCL all_gather of the ranks' local records returns."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    world = 2
    procs = [mpc.Process(target=_ipc_worker, args=(r, world, port, 3000, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert ok, (rank, err)
