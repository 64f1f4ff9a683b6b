"""CPU tests of the chain file format: emcee's HDF5 backend layout of the reference
(gpyrn/meanfield.py:1253-1255) written and read by gpyrn_b200.h5chain without h5py.

The reader is first pinned to a GENUINE HDF5 file (written by MATLAB 7.4 through libhdf5, shipped with scipy's test data)
whose content is known from its MATLAB-5 twin; the writer is then checked through that reader and byte-wise against the
encodings the genuine file holds (datatype / dataspace messages, heap free list, B-tree node, symbol node)."""
import os
import struct

import numpy as np
import pytest

from gpyrn_b200 import h5chain
from gpyrn_b200.h5chain import H5Reader, H5Writer, read_chain, write_chain
from gpyrn_b200.sampler import EnsembleSampler, HDFBackend, NpzBackend, backend_for


def _genuine():
    import scipy.io
    d = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data")
    h5, v5 = os.path.join(d, "testhdf5_7.4_GLNX86.mat"), os.path.join(d, "testdouble_7.4_GLNX86.mat")
    if not (os.path.exists(h5) and os.path.exists(v5)):
        pytest.skip("scipy's HDF5 sample file is not installed")
    return h5, v5


def test_reader_on_a_genuine_libhdf5_file():
    import scipy.io
    h5, v5 = _genuine()
    r = H5Reader(h5)
    assert r.base == 512                                    # MATLAB's 512-byte user block
    assert list(r.root.children) == ["testdouble"]
    node = r["testdouble"]
    want = scipy.io.loadmat(v5)["testdouble"]               # (1, 9) in MATLAB = (9, 1) in HDF5's row-major order
    assert node.data.dtype == np.float64 and node.data.shape == (9, 1)
    assert np.array_equal(node.data.ravel(), want.ravel())
    assert node.attrs == {"MATLAB_class": "double"}


def test_writer_encodings_equal_the_genuine_file():
    h5, _ = _genuine()
    r = H5Reader(h5)
    sb = r.b[512:608]
    root_hdr, bt, hp = struct.unpack_from('<Q', sb, 64)[0], *struct.unpack_from('<QQ', sb, 80)
    child = struct.unpack('<Q', r._at(bt + 32, 8))[0]
    ds_hdr = struct.unpack('<QQ', r._at(child + 8, 16))[1]
    msgs = dict(r._messages(ds_hdr))
    assert msgs[3][:20] == h5chain._dtype_msg('f8')                              # IEEE double, little endian
    assert msgs[1] == h5chain._space_msg((9, 1))
    assert msgs[5] == struct.pack('<BBBBI', 1, 2, 2, 1, 0)                       # the fill-value message the writer emits
    mine_attr = h5chain._attribute("MATLAB_class", "doubl")[8:]                  # a string of size 6 (5 characters + NUL)
    assert msgs[12][:40] == mine_attr[:40] and msgs[12][40:46] == b"double"      # sizes, name, string type, scalar space
    # what the writer produces for the same content, structure by structure
    w = H5Writer()
    w.dataset(None, "testdouble", np.arange(9.0).reshape(9, 1))
    mine = H5Reader(w.tobytes())
    msb = mine.b[:96]
    assert msb[:16] == sb[:16] and msb[16:20] == sb[16:20]                       # signature, versions, sizes, K values
    assert struct.unpack_from('<I', msb, 72)[0] == struct.unpack_from('<I', sb, 72)[0] == 1   # cached B-tree / heap
    mbt, mhp = struct.unpack_from('<QQ', msb, 80)
    assert mine._at(mbt, 8) == r._at(bt, 8)                                      # TREE, group node, leaf, one child
    assert mine._at(mbt + 8, 16) == r._at(bt + 8, 16) == b'\xff' * 16            # no siblings
    assert mine._at(mbt + 24, 8) == r._at(bt + 24, 8) and mine._at(mbt + 40, 8) == r._at(bt + 40, 8)   # keys 0 and 8
    mchild = struct.unpack('<Q', mine._at(mbt + 32, 8))[0]
    assert mine._at(mchild, 16) == r._at(child, 16)                              # SNOD, version, one symbol at name offset 8
    for rd, heap in ((r, hp), (mine, mhp)):                                      # heap: names, then ONE free block
        _, _, size, free, data = struct.unpack('<4sB3xQQQ', rd._at(heap, 32))
        seg = rd._at(data, size)
        assert seg[:8] == bytes(8) and seg[8:19] == b"testdouble\0" and free == 24
        nxt, fsize = struct.unpack_from('<QQ', seg, free)
        assert nxt == 1 and free + fsize == size                                 # 1 = end of the free list
    assert len(mine.b) == struct.unpack_from('<Q', msb, 40)[0]                   # end-of-file address = file size


def test_chain_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    chain, lp = rng.normal(size=(17, 6, 3)), rng.normal(size=(17, 6))
    acc, blobs = rng.integers(0, 17, 6).astype(float), rng.normal(size=(17, 6, 1))
    fn = str(tmp_path / "gprn.h5")
    write_chain(fn, chain, lp, acc, blobs)
    d = read_chain(fn)
    assert d["version"] and d["nwalkers"] == 6 and d["ndim"] == 3 and d["iteration"] == 17 and d["has_blobs"]
    assert d["has_blobs"].dtype == np.bool_                                      # h5py's boolean: enum {FALSE, TRUE} on int8
    assert np.array_equal(d["chain"], chain) and np.array_equal(d["log_prob"], lp)
    assert np.array_equal(d["accepted"], acc) and np.array_equal(d["blobs"], blobs[:, :, 0])
    write_chain(fn, chain, lp, acc, np.dstack([blobs, blobs]))                   # several blobs per walker
    assert read_chain(fn)["blobs"].shape == (17, 6, 2)
    write_chain(fn, np.zeros((0, 6, 3)), np.zeros((0, 6)), np.zeros(6))          # what the reference's be.reset() leaves
    d = read_chain(fn)
    assert d["chain"].shape == (0, 6, 3) and d["iteration"] == 0 and not d["has_blobs"] and "blobs" not in d
    with open(fn, "rb") as f:
        raw = f.read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and len(raw) % 8 == 0
    with pytest.raises(ValueError):
        H5Reader(b"not an hdf5 file" * 8)


def test_groups_with_many_members_and_nesting():
    w = H5Writer()
    for i in range(40):                                                          # five full symbol nodes
        w.dataset(None, f"d{i:02d}", np.full(3, i))
    g = w.group("a")
    gg = w.group("b", g)
    w.attr(gg, "flag", np.bool_(True)); w.attr(gg, "v", np.arange(4.0)); w.attr(gg, "n", 7); w.attr(gg, "s", "text")
    w.dataset(gg, "x", np.arange(5))
    r = H5Reader(w.tobytes())
    assert sorted(r.root.children) == ["a"] + [f"d{i:02d}" for i in range(40)]
    assert all(np.array_equal(r[f"d{i:02d}"].data, np.full(3, i)) for i in range(40))
    assert r["a/b"].attrs["flag"] == True and r["a/b"].attrs["n"] == 7 and r["a/b"].attrs["s"] == "text"  # noqa: E712
    assert np.array_equal(r["a/b"].attrs["v"], np.arange(4.0)) and np.array_equal(r["a/b/x"].data, np.arange(5))
    assert "b" in r["a"] and "c" not in r["a"]
    w = H5Writer()
    w.dataset(None, "c", np.array(["x"]))                                        # no string datasets in this subset
    with pytest.raises(TypeError):
        w.tobytes()


def test_hdf_backend_of_the_sampler(tmp_path):
    fn = str(tmp_path / "gprn.h5")
    assert isinstance(backend_for(fn), HDFBackend) and type(backend_for(str(tmp_path / "c.npz"))) is NpzBackend
    s = EnsembleSampler(4, 2, lambda x: np.column_stack([-0.5 * np.sum(x ** 2, axis=1), x[:, 0]]), seed=0,
                        backend=HDFBackend(fn, every=7))
    d = HDFBackend.load(fn)                                                      # reset: an empty chain of the right shape
    assert d["iteration"] == 0 and d["chain"].shape == (0, 4, 2) and d["log_prob"].shape == (0, 4)
    s.run_mcmc(np.random.default_rng(2).standard_normal((4, 2)), 30)
    d = HDFBackend.load(fn)
    assert d["iteration"] == 30 and d["nwalkers"] == 4 and d["ndim"] == 2 and d["has_blobs"]
    assert np.array_equal(d["chain"], s.get_chain()) and np.array_equal(d["log_prob"], s.get_log_prob())
    assert np.array_equal(d["blobs"], s.get_blobs()[:, :, 0]) and np.array_equal(d["accepted"], s.naccepted)


def test_round_trip_of_random_trees():
    """Property test: random group trees with float / integer / boolean datasets and attributes of random shapes come
    back from the reader exactly as written (names in byte order, empty arrays, scalars, nesting three levels deep)."""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_0123456789", min_size=1, max_size=12)
    arrays = st.one_of(
        hnp.arrays(np.float64, hnp.array_shapes(min_dims=0, max_dims=3, min_side=0, max_side=5),
                   elements=st.floats(allow_nan=False, width=64)),
        hnp.arrays(np.int64, hnp.array_shapes(min_dims=1, max_dims=2, min_side=0, max_side=6)),
        hnp.arrays(np.bool_, hnp.array_shapes(min_dims=0, max_dims=1, min_side=1, max_side=4)))
    attrs = st.dictionaries(names, st.one_of(arrays, st.text(alphabet="abcXYZ .-", max_size=9)), max_size=3)
    leaf = st.fixed_dictionaries({"attrs": attrs, "data": st.dictionaries(names, arrays, max_size=10)})
    tree = st.recursive(leaf, lambda kids: st.fixed_dictionaries(
        {"attrs": attrs, "data": st.dictionaries(names, arrays, max_size=4), "groups": st.dictionaries(names, kids, max_size=3)}),
        max_leaves=6)

    def put(w, g, node):
        for k, v in node["attrs"].items():
            w.attr(g, k, v)
        for k, v in node["data"].items():
            w.dataset(g, k, v)
        for k, v in node.get("groups", {}).items():
            if k not in node["data"]:
                put(w, w.group(k, g if g is not None else w.root), v)

    def same(a, b):
        a, b = np.asarray(a), np.asarray(b)
        return a.shape == b.shape and a.dtype.kind == b.dtype.kind and np.array_equal(a, b)

    def check(r, node):
        for k, v in node["attrs"].items():
            assert (r.attrs[k] == v) if isinstance(v, str) else same(r.attrs[k], v)
        for k, v in node["data"].items():
            assert r.children[k].children is None and same(r.children[k].data, v)
        sub = {k: v for k, v in node.get("groups", {}).items() if k not in node["data"]}
        assert sorted(r.children or {}) == sorted(set(node["data"]) | set(sub))
        for k, v in sub.items():
            check(r.children[k], v)

    @settings(max_examples=60, deadline=None)
    @given(tree)
    def run(node):
        w = H5Writer()
        put(w, None, node)
        check(H5Reader(w.tobytes()).root, node)

    run()
