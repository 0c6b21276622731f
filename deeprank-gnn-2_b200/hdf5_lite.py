"""Minimal read-only HDF5 reader (no h5py / libhdf5 in this image).

DeepRank2 graph files are written by ``Graph.write_to_hdf5`` (reference
``deeprank2/utils/graph.py:210-264``) with bare ``create_dataset`` calls: superblock
v0, "old style" groups (symbol table + v1 B-tree + local heap), v1 object headers
and contiguous, unfiltered datasets of fixed-width ints/floats/strings.  That subset
-- and nothing else -- is implemented here; anything outside it (chunked / filtered
grids ``utils/grid.py:326-333``, variable-length strings, new-style link messages)
raises ``NotImplementedError`` instead of guessing.

The API mirrors the slice of h5py that ``GraphDataset`` uses: ``File(path)`` is a
context manager and a mapping of groups; ``group[name]`` / ``"a/b/c"`` paths,
``name in group``, ``group.keys()`` and ``dataset[()]`` -> numpy.
"""
from __future__ import annotations

import struct

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        if buf[:8] != _SIGNATURE:
            raise ValueError("not an HDF5 file (bad signature)")
        version = buf[8]
        if version not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {version} (only 0/1 = h5py 'earliest' layout)")
        self.size_o = buf[13]
        self.size_l = buf[14]
        if self.size_o != 8 or self.size_l != 8:
            raise NotImplementedError("only 8-byte offsets/lengths")
        pos = 24 + (4 if version == 1 else 0)
        self.base, _free, _eof, _drv = struct.unpack_from("<4Q", buf, pos)
        pos += 32
        self.root_entry = self._symbol_entry(pos)

    # -- primitives
    def u(self, pos: int, n: int) -> int:
        return int.from_bytes(self.buf[pos : pos + n], "little")

    def _symbol_entry(self, pos: int):
        name_off, header_addr, cache_type = struct.unpack_from("<QQI", self.buf, pos)
        btree = heap = None
        if cache_type == 1:
            btree, heap = struct.unpack_from("<QQ", self.buf, pos + 24)
        return name_off, header_addr, cache_type, btree, heap

    # -- object headers
    def messages(self, addr: int):
        """Yield (type, flags, payload_offset, size) of a version-1 object header incl. continuations."""
        buf = self.buf
        if buf[addr] != 1:
            raise NotImplementedError(f"object header version {buf[addr]} (only v1)")
        nmsg = self.u(addr + 2, 2)
        hdr_size = self.u(addr + 8, 4)
        blocks = [(addr + 16, hdr_size)]
        seen = 0
        while blocks and seen < nmsg:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, pos)
                payload = pos + 8
                seen += 1
                if mtype == 0x0010:
                    c_off, c_len = struct.unpack_from("<QQ", buf, payload)
                    blocks.append((c_off + self.base, c_len))
                else:
                    yield mtype, mflags, payload, msize
                pos = payload + msize

    # -- groups
    def group_links(self, btree: int, heap: int) -> dict[str, int]:
        buf = self.buf
        if buf[heap : heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        heap_data = self.u(heap + 24, 8) + self.base
        out: dict[str, int] = {}

        def walk(node: int):
            if buf[node : node + 4] != b"TREE":
                raise ValueError("bad B-tree signature")
            level = buf[node + 5]
            used = self.u(node + 6, 2)
            pos = node + 24
            for i in range(used):
                child = self.u(pos + 8 + i * 16, 8) + self.base
                if level > 0:
                    walk(child)
                else:
                    self._snod(child, heap_data, out)

        walk(btree + self.base)
        return out

    def _snod(self, addr: int, heap_data: int, out: dict[str, int]):
        buf = self.buf
        if buf[addr : addr + 4] != b"SNOD":
            raise ValueError("bad symbol node signature")
        count = self.u(addr + 6, 2)
        for i in range(count):
            name_off, header_addr, _ct, _b, _h = self._symbol_entry(addr + 8 + i * 40)
            start = heap_data + name_off
            end = buf.index(b"\x00", start)
            out[buf[start:end].decode("utf-8")] = header_addr + self.base


def _parse_dtype(buf: bytes, pos: int):
    cls_ver = buf[pos]
    cls = cls_ver & 0x0F
    bits0 = buf[pos + 1]
    size = int.from_bytes(buf[pos + 4 : pos + 8], "little")
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:
        signed = bool(bits0 & 0x08)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
    if cls == 1:
        return np.dtype(f"{order}f{size}")
    if cls == 3:
        return np.dtype(f"S{size}")
    if cls == 8:  # enumeration (h5py stores numpy bool as an int8 enum): values are the base type's
        return _parse_dtype(buf, pos + 8)
    raise NotImplementedError(f"HDF5 datatype class {cls}")


class Dataset:
    def __init__(self, rd: _Reader, addr: int, name: str):
        self._rd = rd
        self.name = name
        shape = None
        dtype = None
        data_addr = None
        data_size = None
        compact = None
        for mtype, _fl, p, _sz in rd.messages(addr):
            buf = rd.buf
            if mtype == 0x0001:
                ver, rank = buf[p], buf[p + 1]
                dims_at = p + (8 if ver == 1 else 4)
                shape = tuple(rd.u(dims_at + 8 * i, 8) for i in range(rank))
            elif mtype == 0x0003:
                dtype = _parse_dtype(buf, p)
            elif mtype == 0x0008:
                ver = buf[p]
                if ver != 3:
                    raise NotImplementedError(f"data layout message version {ver}")
                lclass = buf[p + 1]
                if lclass == 1:
                    data_addr, data_size = struct.unpack_from("<QQ", buf, p + 2)
                elif lclass == 0:
                    csize = rd.u(p + 2, 2)
                    compact = bytes(buf[p + 4 : p + 4 + csize])
                else:
                    raise NotImplementedError(f"{name}: chunked/filtered datasets are outside the graph layout")
            elif mtype == 0x000B:
                raise NotImplementedError(f"{name}: filtered dataset")
        if shape is None or dtype is None:
            raise ValueError(f"{name}: incomplete dataset header")
        self.shape = shape
        self.dtype = dtype
        self._addr = data_addr
        self._size = data_size
        self._compact = compact

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def read(self) -> np.ndarray:
        count = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = count * self.dtype.itemsize
        if self._compact is not None:
            raw = self._compact[:nbytes]
        elif self._addr is None or self._addr == _UNDEF:
            raw = bytes(nbytes)  # never written: fill value 0
        else:
            start = self._addr + self._rd.base
            raw = self._rd.buf[start : start + nbytes]
        arr = np.frombuffer(raw, dtype=self.dtype, count=count).reshape(self.shape)
        arr = arr.astype(self.dtype.newbyteorder("="), copy=True)
        return arr

    def __getitem__(self, key):
        arr = self.read()
        if key == ():
            return arr[()] if arr.ndim == 0 else arr
        return arr[key]

    def __len__(self):
        return self.shape[0]


class Group:
    def __init__(self, rd: _Reader, addr: int, name: str, links: dict[str, int] | None = None):
        self._rd = rd
        self._addr = addr
        self.name = name
        self._links = links

    def _load(self) -> dict[str, int]:
        if self._links is None:
            btree = heap = None
            for mtype, _fl, p, _sz in self._rd.messages(self._addr):
                if mtype == 0x0011:
                    btree, heap = struct.unpack_from("<QQ", self._rd.buf, p)
                elif mtype in (0x0002, 0x0006):
                    raise NotImplementedError(f"{self.name}: new-style (link message) groups are not supported")
            self._links = {} if btree is None else self._rd.group_links(btree, heap + self._rd.base)
        return self._links

    def keys(self):
        return list(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, path: str) -> bool:
        try:
            self[path]
        except KeyError:
            return False
        return True

    def _is_group(self, addr: int) -> bool:
        return any(mtype == 0x0011 for mtype, *_ in self._rd.messages(addr))

    def __getitem__(self, path: str):
        node = self
        parts = [p for p in path.split("/") if p]
        for i, part in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(f"{path!r}: no member {part!r} in {node.name!r}")
            addr = links[part]
            child_name = f"{node.name.rstrip('/')}/{part}"
            node = Group(self._rd, addr, child_name) if self._is_group(addr) else Dataset(self._rd, addr, child_name)
        return node

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class File(Group):
    """``with File(path) as f5: f5[entry]['node_features/_position'][()]``."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise NotImplementedError("hdf5_lite is read-only")
        with open(path, "rb") as fh:
            buf = fh.read()
        rd = _Reader(buf)
        _name_off, header_addr, cache_type, btree, heap = rd.root_entry
        links = rd.group_links(btree, heap + rd.base) if cache_type == 1 else None
        super().__init__(rd, header_addr + rd.base, "/", links)
        self.filename = path

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass
