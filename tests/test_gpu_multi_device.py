"""The C ABI with more than one device: sharding of a host batch inside the library, device-pointer calls on a
second device, and concurrent callers on different devices (eb200_init(devices[], ndev, flags),
thread-safe).  Needs two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from elliptic_b200 import _native as nat
    lib = nat.init_devices([0, 1])
    assert lib.eb200_device_count() >= 2
    return lib


def test_host_batch_is_sharded_over_both_devices(two):
    import benchdata
    from elliptic_b200.ec import EC
    ds = benchdata.gen_secp256k1_verify(1 << 17, seed=0xE1110777, cache_dir="/tmp/eb200_cache")
    for fmt in ("pageable", "pinned"):
        arrs = {k: ds[k] for k in ("e", "r", "s", "pub")}
        if fmt == "pinned":
            import torch
            arrs = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in arrs.items()}
        st = EC("secp256k1").verify_batch_packed(arrs["e"], arrs["r"], arrs["s"], arrs["pub"])
        assert np.array_equal(st, ds["expected"]), fmt
    # an odd size: blocks are cut at multiples of 128 items
    m = (1 << 16) + 4099
    st = EC("secp256k1").verify_batch_packed(ds["e"][:m], ds["r"][:m], ds["s"][:m], ds["pub"][:m])
    assert np.array_equal(st, ds["expected"][:m])


def test_device_pointer_calls_follow_the_pointer(two):
    import torch
    import benchdata
    from elliptic_b200 import _native as nat
    lib = two
    ds = benchdata.gen_secp256k1_verify(1 << 14, seed=0xE1110778, cache_dir="/tmp/eb200_cache")
    for g in (1, 0):
        dev = torch.device("cuda", g)
        d = {k: torch.from_numpy(ds[k]).to(dev) for k in ("e", "r", "s", "pub")}
        n = 1 << 14
        st = torch.empty(n, dtype=torch.uint8, device=dev)
        ws = torch.empty(lib.eb200_ecdsa_verify_workspace_bytes(1, n), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            nat.check(lib.eb200_ecdsa_verify_batch_dev(1, n, d["e"].data_ptr(), d["r"].data_ptr(), d["s"].data_ptr(), d["pub"].data_ptr(),
                                                       0, st.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
            torch.cuda.synchronize(dev)
        assert np.array_equal(st.cpu().numpy(), ds["expected"])
        assert nat.last_timing()["main_kernel_ms"] > 0


def test_two_threads_two_devices(two):
    """Concurrent host-pointer callers: small batches take one device each (rotating), so two threads run on two
    devices at once; every call must still return its own results."""
    import benchdata
    from elliptic_b200.ec import EC
    from elliptic_b200.eddsa import EDDSA
    ds = benchdata.gen_secp256k1_verify(1 << 13, seed=0xE1110779, cache_dir="/tmp/eb200_cache")
    de = benchdata.gen_ed25519_verify(1 << 13, seed=0xE111077A, cache_dir="/tmp/eb200_cache")
    errs = []

    def worker(kind, reps):
        try:
            for _ in range(reps):
                if kind == "ecdsa":
                    st = EC("secp256k1").verify_batch_packed(ds["e"], ds["r"], ds["s"], ds["pub"])
                    assert np.array_equal(st, ds["expected"])
                else:
                    st = EDDSA().verify_batch_packed(de["R"], de["S"], de["A"], de["h"])
                    assert np.array_equal(st, de["expected"])
        except Exception as ex:                   # noqa: BLE001
            errs.append(repr(ex))
    ths = [threading.Thread(target=worker, args=(k, 12)) for k in ("ecdsa", "eddsa", "ecdsa", "eddsa")]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
