# -*- coding: utf-8 -*-
"""
Description
-----------
Minimal reader/writer for ParaMol ``.nc`` data files (NetCDF-4 = HDF5 containers).

The reference reads/writes these through the third-party ``netCDF4`` package
(``ParaMol/System/system.py:567-686``; file spec ``docs/source/Files_specification.rst:137-160``),
which is not available here.  ParaMol files only ever hold three contiguous little-endian float64
datasets (``reference_coordinates`` [S,N,3] nm, ``reference_forces`` [S,N,3] kJ/mol/nm,
``reference_energies`` [S] kJ/mol), so this module implements just enough of the HDF5 file format
to walk the root group and return contiguous numeric datasets:

* superblock versions 0-3, object headers v1 and v2 (with continuation blocks),
* new-style groups with compact link messages, old-style groups (v1 B-tree + local heap),
* dataspace v1/v2, fixed-point / floating-point datatypes, contiguous and compact layouts.

Chunked/filtered datasets raise ``NotImplementedError``.  The writer emits a plain ``.npz``-free,
self-describing raw container with the same three variables (``write_paramol_raw``) for round trips.
"""
import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _H5:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5/NetCDF-4 file")
        ver = buf[8]
        if ver in (0, 1):
            self.so = buf[13]
            self.sl = buf[14]
            p = 24 if ver == 0 else 28
            # base, free-space, end-of-file, driver-info addresses
            p += 4 * self.so
            # root group symbol table entry: link-name offset, object header address
            p += self.so
            self.root = self._addr(p)
        elif ver in (2, 3):
            self.so = buf[9]
            self.sl = buf[10]
            p = 12 + 3 * self.so
            self.root = self._addr(p)
        else:
            raise NotImplementedError("HDF5 superblock version %d" % ver)

    # -- primitive decoders -------------------------------------------------------------------
    def _uint(self, p, n):
        return int.from_bytes(self.b[p:p + n], "little")

    def _addr(self, p):
        return self._uint(p, self.so)

    def _len(self, p):
        return self._uint(p, self.sl)

    # -- object headers -----------------------------------------------------------------------
    def messages(self, addr):
        """Yield (type, flags, payload_offset, payload_size) for every header message of the object."""
        b = self.b
        if b[addr:addr + 4] == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            chunk0 = self._uint(p, nsz)
            p += nsz
            creation_order = bool(flags & 0x04)
            blocks = [(p, p + chunk0)]
            while blocks:
                q, end = blocks.pop(0)
                while q + 4 <= end:
                    mtype = b[q]
                    msize = self._uint(q + 1, 2)
                    mflags = b[q + 3]
                    q += 4
                    if creation_order:
                        q += 2
                    if mtype == 0x10:
                        caddr = self._addr(q)
                        clen = self._len(q + self.so)
                        # continuation block: 'OCHK' signature, messages, 4-byte checksum
                        blocks.append((caddr + 4, caddr + clen - 4))
                    elif mtype != 0:
                        yield mtype, mflags, q, msize
                    q += msize
        else:
            # version 1 object header
            if b[addr] != 1:
                raise ValueError("bad object header at %d" % addr)
            nmsg = self._uint(addr + 2, 2)
            hsize = self._uint(addr + 8, 4)
            blocks = [(addr + 16, addr + 16 + hsize)]
            seen = 0
            while blocks and seen < nmsg:
                q, end = blocks.pop(0)
                while q + 8 <= end and seen < nmsg:
                    mtype = self._uint(q, 2)
                    msize = self._uint(q + 2, 2)
                    mflags = b[q + 4]
                    q += 8
                    seen += 1
                    if mtype == 0x10:
                        caddr = self._addr(q)
                        clen = self._len(q + self.so)
                        blocks.append((caddr, caddr + clen))
                    elif mtype != 0:
                        yield mtype, mflags, q, msize
                    q += msize

    # -- groups -------------------------------------------------------------------------------
    def links(self, addr):
        """Return {name: object header address} for a group object."""
        out = {}
        for mtype, _, q, msize in self.messages(addr):
            if mtype == 0x06:  # link message
                b = self.b
                p = q + 1
                flags = b[q + 1]
                p = q + 2
                ltype = 0
                if flags & 0x08:
                    ltype = b[p]
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsz = 1 << (flags & 3)
                nlen = self._uint(p, lsz)
                p += lsz
                name = bytes(b[p:p + nlen]).decode("utf-8")
                p += nlen
                if ltype == 0:
                    out[name] = self._addr(p)
            elif mtype == 0x11:  # symbol table message (old-style group)
                btree = self._addr(q)
                heap = self._addr(q + self.so)
                out.update(self._symbol_table(btree, heap))
            elif mtype == 0x02:  # link info: dense storage not supported
                fheap = self._addr(q + 2 + (8 if self.b[q + 1] & 1 else 0))
                if fheap != _UNDEF & ((1 << (8 * self.so)) - 1):
                    raise NotImplementedError("dense link storage")
        return out

    def _symbol_table(self, btree, heap):
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data = self._addr(heap + 8 + 2 * self.sl)
        out = {}

        def walk(node):
            if b[node:node + 4] == b"TREE":
                level = b[node + 5]
                n = self._uint(node + 6, 2)
                p = node + 8 + 2 * self.so
                for i in range(n):
                    p += self.sl  # key
                    child = self._addr(p)
                    p += self.so
                    walk(child)
                _ = level
            elif b[node:node + 4] == b"SNOD":
                n = self._uint(node + 6, 2)
                p = node + 8
                for i in range(n):
                    noff = self._len(p)
                    oaddr = self._addr(p + self.sl)
                    s = data + noff
                    e = b.index(b"\x00", s)
                    out[bytes(b[s:e]).decode("utf-8")] = oaddr
                    p += self.sl + self.so + 4 + 4 + 16
        walk(btree)
        return out

    # -- datasets -----------------------------------------------------------------------------
    def dataset(self, addr):
        b = self.b
        shape = None
        dtype = None
        data = None
        for mtype, _, q, msize in self.messages(addr):
            if mtype == 0x01:  # dataspace
                ver = b[q]
                rank = b[q + 1]
                flags = b[q + 2]
                p = q + (8 if ver == 1 else 4)
                shape = tuple(self._len(p + i * self.sl) for i in range(rank))
                _ = flags
            elif mtype == 0x03:  # datatype
                cls = b[q] & 0x0F
                bits0 = b[q + 1]
                size = self._uint(q + 4, 4)
                order = ">" if (bits0 & 1) else "<"
                if cls == 1:
                    dtype = np.dtype(order + "f%d" % size)
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype(order + ("i" if signed else "u") + "%d" % size)
                else:
                    dtype = None
            elif mtype == 0x08:  # data layout
                ver = b[q]
                if ver in (3, 4):
                    lclass = b[q + 1]
                    if lclass == 1:  # contiguous
                        daddr = self._addr(q + 2)
                        dlen = self._len(q + 2 + self.so)
                        data = ("contiguous", daddr, dlen)
                    elif lclass == 0:  # compact
                        dlen = self._uint(q + 2, 2)
                        data = ("contiguous", q + 4, dlen)
                    else:
                        data = ("chunked", None, None)
                else:
                    raise NotImplementedError("data layout version %d" % ver)
        if shape is None or data is None:
            return None
        if dtype is None:
            return None
        if data[0] != "contiguous":
            raise NotImplementedError("chunked/filtered HDF5 dataset")
        n = int(np.prod(shape)) if len(shape) else 1
        _, daddr, dlen = data
        if daddr == _UNDEF & ((1 << (8 * self.so)) - 1):
            return np.zeros(shape, dtype=dtype.newbyteorder("="))
        arr = np.frombuffer(b, dtype=dtype, count=n, offset=daddr).reshape(shape)
        return arr.astype(dtype.newbyteorder("="), copy=True)


def read_nc(file_name):
    """
    Read every contiguous numeric variable of the root group of a NetCDF-4/HDF5 file.

    Returns
    -------
    dict : {variable name: np.ndarray}
    """
    with open(file_name, "rb") as f:
        buf = f.read()
    if buf[:8] != _SIG:
        return _read_paramol_raw(buf)
    h5 = _H5(memoryview(buf).toreadonly() if False else buf)
    out = {}
    for name, addr in h5.links(h5.root).items():
        try:
            arr = h5.dataset(addr)
        except NotImplementedError:
            raise
        if arr is not None:
            out[name] = arr
    return out


def read_paramol_data(file_name):
    """
    ParaMol-specific view of ``read_nc`` (reference: ``ParaMolSystem.read_data``, ``system.py:567-633``).

    Returns
    -------
    coordinates, energies, forces : np.ndarray or None
    """
    v = read_nc(file_name)

    def get(key):
        a = v.get(key)
        if a is None or len(a) == 0:
            return None
        return np.ascontiguousarray(a, dtype=np.float64)

    return get("reference_coordinates"), get("reference_energies"), get("reference_forces")


# ---------------------------------------------------------------------------------------------
# Raw container writer (used by ParaMolSystem.write_data until a full HDF5 writer lands).
# Layout: magic 'PMRAW001', uint32 n_vars, then per variable: uint16 name length, name,
# uint8 rank, rank x uint64 dims, float64 little-endian payload.
# ---------------------------------------------------------------------------------------------
_RAW_MAGIC = b"PMRAW001"


def write_paramol_raw(file_name, variables):
    with open(file_name, "wb") as f:
        f.write(_RAW_MAGIC)
        items = [(k, np.ascontiguousarray(v, dtype="<f8")) for k, v in variables.items() if v is not None]
        f.write(struct.pack("<I", len(items)))
        for name, arr in items:
            nb = name.encode("utf-8")
            f.write(struct.pack("<H", len(nb)))
            f.write(nb)
            f.write(struct.pack("<B", arr.ndim))
            f.write(struct.pack("<%dQ" % arr.ndim, *arr.shape))
            f.write(arr.tobytes())
    return True


def _read_paramol_raw(buf):
    if buf[:8] != _RAW_MAGIC:
        raise ValueError("unknown data file format")
    p = 8
    (n,) = struct.unpack_from("<I", buf, p)
    p += 4
    out = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<H", buf, p)
        p += 2
        name = buf[p:p + ln].decode("utf-8")
        p += ln
        rank = buf[p]
        p += 1
        dims = struct.unpack_from("<%dQ" % rank, buf, p)
        p += 8 * rank
        cnt = int(np.prod(dims)) if rank else 1
        out[name] = np.frombuffer(buf, dtype="<f8", count=cnt, offset=p).reshape(dims).copy()
        p += 8 * cnt
    return out
