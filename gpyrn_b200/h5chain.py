"""The on-disk format of the reference's MCMC chain: emcee's ``HDFBackend("gprn.h5")`` (gpyrn/meanfield.py:1253-1255).

Neither ``h5py`` nor ``emcee`` exists in this image, so this module writes and reads the HDF5 *file format* itself --
the subset the backend's layout needs, in the oldest and most widely readable encoding (what libhdf5 1.8 writes with
default settings):

* superblock version 0, 8-byte offsets and lengths;
* groups as symbol tables (version-1 B-tree node ``TREE`` + local heap ``HEAP`` + symbol node ``SNOD``);
* version-1 object headers with dataspace (v1), datatype (v1), fill value (v1), data layout (v3, contiguous) and
  attribute (v1) messages;
* IEEE little-endian float64, int64 / int8, fixed-length ASCII strings and h5py's boolean (``enum {FALSE, TRUE}`` on int8).

The layout inside the file is emcee 3's (``emcee/backends/hdf.py``): a group (default ``"mcmc"``) with the attributes
``version``, ``nwalkers``, ``ndim``, ``has_blobs``, ``iteration`` and the datasets ``accepted`` (nwalkers,), ``chain``
(iteration, nwalkers, ndim), ``log_prob`` (iteration, nwalkers) and, if the log-probability returns them, ``blobs``
(iteration, nwalkers[, nblobs]).  ``h5py.File("gprn.h5")["mcmc"]["chain"][...]`` and
``emcee.backends.HDFBackend("gprn.h5", read_only=True).get_chain()`` are the intended readers.  The datasets are
contiguous (fixed shape): the file is rewritten at every save instead of grown in place, so emcee cannot *append* to it.

Validation available without libhdf5: :class:`H5Reader` is checked (tests/test_h5chain.py) against a genuine HDF5 file
that ships with scipy's test data (written by MATLAB 7.4 through libhdf5: same superblock / group / header encoding,
contiguous datasets, attributes), and then reads back what :class:`H5Writer` produced.  Host-side convenience only:
nothing here is on the device path.
"""
import struct

import numpy as np

__all__ = ['H5Writer', 'H5Reader', 'write_chain', 'read_chain']

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K = 4, 16           # libhdf5 defaults: up to 2*_LEAF_K entries per symbol node
_MSG_DATASPACE, _MSG_DATATYPE, _MSG_FILL, _MSG_LAYOUT, _MSG_ATTR, _MSG_CONT, _MSG_SYMTAB = 1, 3, 5, 8, 12, 16, 17


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


# ---- datatype / dataspace messages ------------------------------------------------------------------------------------
def _dtype_msg(kind, size=0):
    """Datatype message (version 1) of 'f8', 'i8', 'i1', 'bool' (h5py's enum on int8) or 'S' (fixed-length string)."""
    if kind == 'f8':      # class 1; little endian, mantissa normalisation 2 (implied), sign bit 63
        return struct.pack('<B3BI', 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
    if kind in ('i8', 'i1'):   # class 0; little endian, two's complement signed
        n = int(kind[1])
        return struct.pack('<B3BI', 0x10, 0x08, 0x00, 0x00, n) + struct.pack('<HH', 0, 8 * n)
    if kind == 'bool':    # class 8 (enumeration), 2 members, base type int8, names padded to 8 bytes, then the values
        return (struct.pack('<B3BI', 0x18, 0x02, 0x00, 0x00, 1) + _dtype_msg('i1') + _pad8(b'FALSE\0') + _pad8(b'TRUE\0')
                + bytes([0, 1]))
    if kind == 'S':       # class 3; null-terminated, ASCII
        return struct.pack('<B3BI', 0x13, 0x00, 0x00, 0x00, size)
    raise TypeError(f"no HDF5 datatype for {kind!r}")


def _space_msg(shape):
    """Dataspace message, version 1 (no maximum dimensions, no permutation); shape () is a scalar."""
    return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + b''.join(struct.pack('<Q', int(n)) for n in shape)


def _kind_of(a):
    if a.dtype == np.bool_:
        return 'bool', a.astype(np.int8)
    if a.dtype.kind == 'f':
        return 'f8', a.astype('<f8')
    if a.dtype.kind in 'iu':
        return 'i8', a.astype('<i8')
    raise TypeError(f"unsupported array type {a.dtype}")


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _object_header(messages):
    body = b''.join(messages)
    return struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body


def _attribute(name, value):
    """Attribute message (version 1): scalar numbers / booleans, small arrays, or a string."""
    if isinstance(value, (str, bytes)):
        raw = (value.encode('ascii') if isinstance(value, str) else value) + b'\0'
        dt, sp, data = _dtype_msg('S', len(raw)), _space_msg(()), raw
    else:
        kind, a = _kind_of(np.asarray(value))
        dt, sp, data = _dtype_msg(kind), _space_msg(a.shape), a.tobytes()
    nm = name.encode('ascii') + b'\0'
    return _message(_MSG_ATTR, struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp)
                    + data)


# ---- writer -------------------------------------------------------------------------------------------------------
class _Group:
    def __init__(self):
        self.attrs, self.children = {}, {}          # name -> value ; name -> _Group | ndarray


class H5Writer:
    """Collects groups, datasets and attributes, then lays the file out in one pass (:meth:`tobytes`).

    >>> w = H5Writer(); g = w.group("mcmc"); w.attr(g, "ndim", 3); w.dataset(g, "chain", np.zeros((10, 6, 3)))
    >>> w.save("gprn.h5")
    """

    def __init__(self):
        self.root = _Group()
        self._out = bytearray()

    def group(self, name, parent=None):
        parent = parent or self.root
        return parent.children.setdefault(name, _Group())

    def dataset(self, parent, name, array):
        (parent or self.root).children[name] = np.array(array, order='C')       # (ascontiguousarray would make scalars 1-d)

    def attr(self, parent, name, value):
        (parent or self.root).attrs[name] = value

    # -- layout --------------------------------------------------------------------------------------------------
    def _alloc(self, blob):
        self._out += b'\0' * (-len(self._out) % 8)
        addr = len(self._out)
        self._out += blob
        return addr

    def _put_dataset(self, a):
        kind, a = _kind_of(a)
        raw = a.tobytes()
        addr = self._alloc(raw) if raw else _UNDEF
        msgs = [_message(_MSG_DATASPACE, _space_msg(a.shape)), _message(_MSG_DATATYPE, _dtype_msg(kind), flags=1),
                _message(_MSG_FILL, struct.pack('<BBBBI', 1, 2, 2, 1, 0), flags=1),   # late allocation, fill if set, default value
                _message(_MSG_LAYOUT, struct.pack('<BBQQ', 3, 1, addr, len(raw)))]
        return self._alloc(_object_header(msgs))

    def _put_group(self, g):
        """Writes the children, then heap, symbol node(s), B-tree node and the group's object header.  Returns
        (header address, B-tree address, heap address)."""
        if len(g.children) > 2 * _LEAF_K * 2 * _INTERNAL_K:
            raise ValueError("too many members for a one-level group B-tree")
        entries = []
        for name in sorted(g.children, key=lambda s: s.encode('ascii')):
            child = g.children[name]
            if isinstance(child, _Group):
                hdr, bt, hp = self._put_group(child)
                entries.append((name, hdr, 1, struct.pack('<QQ', bt, hp)))
            else:
                entries.append((name, self._put_dataset(child), 0, b'\0' * 16))
        # local heap: the empty name at offset 0, the member names, one free block at the end
        heap, offsets = bytearray(8), []
        for name, *_ in entries:
            offsets.append(len(heap))
            heap += _pad8(name.encode('ascii') + b'\0')
        free_at = len(heap)
        heap += struct.pack('<QQ', 1, 32) + b'\0' * 16                         # next free block: none (1); size 32
        self._out += b'\0' * (-len(self._out) % 8)                              # prefix, data segment right behind it
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_at, len(self._out) + 32) + bytes(heap))
        # symbol nodes of up to 2*_LEAF_K entries each, then one leaf-level B-tree node over them
        keys, nodes = [0], []
        for s in range(0, len(entries), 2 * _LEAF_K):
            part = list(zip(entries, offsets))[s:s + 2 * _LEAF_K]
            body = b''.join(struct.pack('<QQII', off, hdr, cache, 0) + scratch for (_, hdr, cache, scratch), off in part)
            nodes.append(self._alloc(b'SNOD' + struct.pack('<BBH', 1, 0, len(part)) + body.ljust(2 * _LEAF_K * 40, b'\0')))
            keys.append(part[-1][1])
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(nodes), _UNDEF, _UNDEF)
        tree += b''.join(struct.pack('<QQ', k, n) for k, n in zip(keys, nodes)) + struct.pack('<Q', keys[-1])
        bt_addr = self._alloc(tree.ljust(24 + (4 * _INTERNAL_K + 1) * 8, b'\0'))
        msgs = [_message(_MSG_SYMTAB, struct.pack('<QQ', bt_addr, heap_addr))]
        msgs += [_attribute(k, v) for k, v in g.attrs.items()]
        return self._alloc(_object_header(msgs)), bt_addr, heap_addr

    def tobytes(self):
        self._out = bytearray(96)                                                # superblock written last
        hdr, bt, hp = self._put_group(self.root)
        self._out += b'\0' * (-len(self._out) % 8)
        sb = _SIG + struct.pack('<8BHHI', 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, _UNDEF, len(self._out), _UNDEF)
        sb += struct.pack('<QQII', 0, hdr, 1, 0) + struct.pack('<QQ', bt, hp)   # root symbol-table entry, cached addresses
        self._out[:96] = sb
        return bytes(self._out)

    def save(self, filename):
        with open(filename, 'wb') as f:
            f.write(self.tobytes())


# ---- reader -------------------------------------------------------------------------------------------------------
class H5Node:
    """A group (``children``: name -> H5Node) or a dataset (``data``: ndarray); both carry ``attrs``."""

    def __init__(self):
        self.attrs, self.children, self.data = {}, None, None

    def __getitem__(self, name):
        node = self
        for part in name.strip('/').split('/'):
            node = node.children[part]
        return node

    def __contains__(self, name):
        return self.children is not None and name in self.children


class H5Reader:
    """Reader for the same subset (plus an optional user block, object-header continuations and the compact layout):
    enough for files written by :class:`H5Writer` and for plain libhdf5 1.6 / 1.8 files with contiguous datasets."""

    def __init__(self, source):
        self.b = source if isinstance(source, (bytes, bytearray)) else open(source, 'rb').read()
        base = 0
        while self.b[base:base + 8] != _SIG:            # the superblock sits at 0, 512, 1024, ... (user block)
            base = 512 if base == 0 else 2 * base
            if base >= len(self.b):
                raise ValueError("not an HDF5 file")
        sb = self.b[base:base + 96]
        if sb[8] != 0 or sb[13] != 8 or sb[14] != 8:
            raise ValueError("only superblock version 0 with 8-byte offsets and lengths is supported")
        self.base = struct.unpack_from('<Q', sb, 24)[0]           # every address in the file is relative to this one
        self.eof = struct.unpack_from('<Q', sb, 40)[0]
        self.root = self._object(struct.unpack_from('<Q', sb, 64)[0])

    def __getitem__(self, name):
        return self.root[name]

    def _at(self, addr, n):
        a = self.base + addr
        if a + n > len(self.b):
            raise ValueError("address past the end of the file")
        return self.b[a:a + n]

    # -- object headers --------------------------------------------------------------------------------------------
    def _messages(self, addr):
        ver, _, nmsg, _, size = struct.unpack('<BBHII', self._at(addr, 12))
        if ver != 1:
            raise ValueError("only version-1 object headers are supported")
        chunks, out = [(addr + 16, size)], []
        while chunks and len(out) < nmsg:
            pos, left = chunks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack('<HHB', self._at(pos, 5))
                body = self._at(pos + 8, msize)
                if mtype == _MSG_CONT:
                    chunks.append(struct.unpack('<QQ', body[:16]))
                out.append((mtype, body))
                pos, left = pos + 8 + msize, left - 8 - msize
        return out

    def _object(self, addr):
        node, shape, dtype, layout = H5Node(), None, None, None
        for mtype, body in self._messages(addr):
            if mtype == _MSG_SYMTAB:
                node.children = self._group(*struct.unpack('<QQ', body[:16]))
            elif mtype == _MSG_DATASPACE:
                shape = self._space(body)
            elif mtype == _MSG_DATATYPE:
                dtype = self._dtype(body)[0]
            elif mtype == _MSG_LAYOUT:
                layout = body
            elif mtype == _MSG_ATTR:
                name, value = self._attr(body)
                node.attrs[name] = value
        if layout is not None:
            node.data = self._data(layout, shape, dtype)
        return node

    @staticmethod
    def _space(body):
        ver, rank = body[0], body[1]
        if ver == 1:
            return struct.unpack_from(f'<{rank}Q', body, 8)
        if ver == 2:
            return struct.unpack_from(f'<{rank}Q', body, 4)
        raise ValueError("unknown dataspace version")

    @staticmethod
    def _dtype(body):
        """-> (numpy dtype or ('S', n) / 'bool', bytes consumed)."""
        cls, ver = body[0] & 15, body[0] >> 4
        bits = body[1] | body[2] << 8 | body[3] << 16
        size = struct.unpack_from('<I', body, 4)[0]
        end = '>' if bits & 1 else '<'
        if cls == 0:
            return np.dtype(f"{end}{'i' if bits & 8 else 'u'}{size}"), 12
        if cls == 1:
            return np.dtype(f"{end}f{size}"), 20
        if cls == 3:
            return np.dtype(f"S{size}"), 8
        if cls == 8:
            base, used = H5Reader._dtype(body[8:])
            pos, names = 8 + used, []
            for _ in range(bits & 0xFFFF):
                z = body.index(b'\0', pos)
                names.append(body[pos:z])
                pos = z + 1 if ver >= 3 else pos + (z - pos + 8) // 8 * 8
            return (np.dtype(np.bool_) if names == [b'FALSE', b'TRUE'] and base.itemsize == 1 else base), pos + size * len(names)
        raise ValueError(f"datatype class {cls} is not supported")

    def _attr(self, body):
        ver, _, nlen, dlen, slen = struct.unpack_from('<BBHHH', body)
        if ver != 1:
            raise ValueError("only version-1 attribute messages are supported")
        r8 = lambda n: (n + 7) // 8 * 8
        pos = 8
        name = body[pos:pos + nlen].split(b'\0')[0].decode('ascii'); pos += r8(nlen)
        dtype = self._dtype(body[pos:pos + dlen])[0]; pos += r8(dlen)
        shape = self._space(body[pos:pos + slen]); pos += r8(slen)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        a = np.frombuffer(body, dtype=dtype, count=n, offset=pos).reshape(shape)
        if a.dtype.kind == 'S':
            return name, a.reshape(-1)[0].split(b'\0')[0].decode('ascii') if not shape else a
        return name, (a[()] if not shape else a.copy())

    def _data(self, layout, shape, dtype):
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] in (1, 2) and layout[2] == 1:          # libhdf5 1.6 and older: version, rank + 1, class, address
            addr = struct.unpack_from('<Q', layout, 8)[0]
            return np.frombuffer(self._at(addr, n * dtype.itemsize), dtype=dtype, count=n).reshape(shape).copy()
        ver, cls = layout[0], layout[1]
        if ver != 3 or cls not in (0, 1):
            raise ValueError("only compact and contiguous layouts are supported")
        if cls == 0:
            size = struct.unpack_from('<H', layout, 2)[0]
            raw = layout[4:4 + size]
        else:
            addr, size = struct.unpack_from('<QQ', layout, 2)
            raw = b'' if addr == _UNDEF or n == 0 else self._at(addr, n * dtype.itemsize)
        if not raw:
            return np.zeros(shape, dtype=dtype)
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()

    # -- groups --------------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        sig, _, size, _, data = struct.unpack('<4sB3xQQQ', self._at(heap_addr, 32))
        if sig != b'HEAP':
            raise ValueError("bad local heap")
        seg = self._at(data, size)
        return seg[off:seg.index(b'\0', off)].decode('ascii')

    def _group(self, bt_addr, heap_addr):
        out = {}
        sig, ntype, level, used = struct.unpack('<4sBBH', self._at(bt_addr, 8))
        if sig != b'TREE' or ntype != 0:
            raise ValueError("bad group B-tree node")
        for i in range(used):
            child = struct.unpack('<Q', self._at(bt_addr + 24 + 8 + 16 * i, 8))[0]
            if level > 0:
                out.update(self._group(child, heap_addr))
                continue
            sig, _, _, nsym = struct.unpack('<4sBBH', self._at(child, 8))
            if sig != b'SNOD':
                raise ValueError("bad symbol node")
            for j in range(nsym):
                off, hdr = struct.unpack('<QQ', self._at(child + 8 + 40 * j, 16))
                out[self._heap_name(heap_addr, off)] = self._object(hdr)
        return out


# ---- emcee's layout -----------------------------------------------------------------------------------------------------
def write_chain(filename, chain, log_prob, accepted, blobs=None, name="mcmc", version="3.1.6"):
    """One chain in emcee's ``HDFBackend`` layout.  ``chain`` (iteration, nwalkers, ndim), ``log_prob`` (iteration,
    nwalkers), ``accepted`` (nwalkers,), ``blobs`` (iteration, nwalkers[, nblobs]) or None."""
    chain = np.asarray(chain, dtype=float)
    it, nwalkers, ndim = chain.shape
    w = H5Writer()
    g = w.group(name)
    for k, v in (("version", version), ("nwalkers", nwalkers), ("ndim", ndim), ("has_blobs", np.bool_(blobs is not None)),
                 ("iteration", it)):
        w.attr(g, k, v)
    w.dataset(g, "accepted", np.asarray(accepted, dtype=float).reshape(nwalkers))
    w.dataset(g, "chain", chain)
    w.dataset(g, "log_prob", np.asarray(log_prob, dtype=float).reshape(it, nwalkers))
    if blobs is not None:
        blobs = np.asarray(blobs, dtype=float).reshape(it, nwalkers, -1)
        w.dataset(g, "blobs", blobs[:, :, 0] if blobs.shape[2] == 1 else blobs)       # one blob per walker: (it, nwalkers)
    w.save(filename)


def read_chain(filename, name="mcmc"):
    """-> dict with the group's attributes and datasets (``chain``, ``log_prob``, ``accepted``, ``blobs`` if present)."""
    g = H5Reader(filename)[name]
    out = dict(g.attrs)
    out.update({k: v.data for k, v in g.children.items() if v.data is not None})
    return out
