

def test_members_of_a_large_pool_is_the_same_list_on_any_number_of_host_threads(hostemu_lib, monkeypatch):
    """Members() of >= 128 Ki members splits the id range over persistent host threads (count, scan, fill):
    same list as the single-threaded walk, including a caller buffer smaller than the list."""
    import ctypes as C
    import numpy as np
    from consul_b200._lib import GsimMember
    from consul_b200.pool import Pool, lan_config
    n = 150_000
    p = Pool(lan_config(hostemu_lib, capacity=n + 8, n_initial=n, seed=77), hostemu_lib)
    p.crash_fraction(200_000, 5)
    p.step(3)
    p.leave(123)
    p.step(40)
    lists = []
    for threads in ("1", "3", "8", "1"):
        monkeypatch.setenv("GSIM_MEMBERS_THREADS", threads)
        buf = (GsimMember * (n + 8))()
        k = C.c_size_t()
        assert hostemu_lib.gsim_members(p.h, 7, buf, n + 8, C.byref(k)) == 0
        lists.append(np.frombuffer(buf, dtype=np.uint32).reshape(-1, 4)[:k.value].copy())
        small = (GsimMember * 1000)()
        k2 = C.c_size_t()
        assert hostemu_lib.gsim_members(p.h, 7, small, 1000, C.byref(k2)) == 0 and k2.value == k.value
        assert (np.frombuffer(small, dtype=np.uint32).reshape(-1, 4) == lists[-1][:1000]).all()
    for other in lists[1:]:
        assert lists[0].shape == other.shape and (lists[0] == other).all()
    assert lists[0].shape[0] == n and (lists[0][:, 0] == np.arange(n)).all()
    p.close()
