"""Minimal pure-Python HDF5 reader (TEST INFRASTRUCTURE, not product code).

No HDF5 library exists in this image, but the reference's golden vectors are HDF5
files written by h5py with the "earliest" format: superblock v0, v1 object headers,
symbol-table groups (v1 B-tree + local heap), contiguous / compact / chunked(+deflate,
+shuffle) datasets of fixed-point and IEEE float types, and variable-length string
attributes stored in a global heap.  This module decodes exactly that subset so that
`oracle/make_golden.py` can turn the reference fixtures
(`/root/reference/simetuc/test/**.hdf5`) into committed `.npz` files under
`tests/golden/`.  The file layout decoded here is the one written by
`lattice.py:527-548` and `simulations.py:227-241` of the reference.
"""
import struct
import zlib

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    """Read-only view of an old-format HDF5 file: `f['name']` -> ndarray, `f.attrs`."""

    def __init__(self, path):
        with open(path, 'rb') as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != _SIG:
            raise ValueError('not an HDF5 file: %s' % path)
        ver = b[8]
        if ver not in (0, 1):
            raise ValueError('unsupported superblock version %d' % ver)
        self.O = b[13]
        self.L = b[14]
        if self.O != 8 or self.L != 8:
            raise ValueError('only 8-byte offsets/lengths supported')
        pos = 24 if ver == 0 else 28
        self.base, _fs, _eof, _drv = struct.unpack_from('<4Q', b, pos)
        pos += 32
        # root group symbol table entry
        _name_off, root_hdr, cache, _r = struct.unpack_from('<QQII', b, pos)
        self.root_msgs = self._read_object_header(root_hdr)
        self.links = self._group_links(self.root_msgs)
        self.attrs = self._attributes(self.root_msgs)

    # ------------------------------------------------------------------ object headers
    def _read_object_header(self, addr):
        b = self.buf
        ver, _r, nmsg, _refc, hsize = struct.unpack_from('<BBHII', b, addr)
        if ver != 1:
            raise ValueError('only v1 object headers supported (got %d)' % ver)
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from('<HHB', b, pos)
                data = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:       # continuation
                    coff, clen = struct.unpack_from('<QQ', data, 0)
                    blocks.append((coff, clen))
                msgs.append((mtype, mflags, data))
        return msgs

    def _group_links(self, msgs):
        links = {}
        for mtype, _f, data in msgs:
            if mtype == 0x11:           # symbol table message
                btree, heap = struct.unpack_from('<QQ', data, 0)
                heap_data = self._local_heap(heap)
                self._walk_group_btree(btree, heap_data, links)
        return links

    def _local_heap(self, addr):
        b = self.buf
        assert b[addr:addr + 4] == b'HEAP'
        size, _free, daddr = struct.unpack_from('<QQQ', b, addr + 8)
        return b[daddr:daddr + size]

    def _walk_group_btree(self, addr, heap, links):
        b = self.buf
        assert b[addr:addr + 4] == b'TREE', 'bad group btree'
        ntype, level, nused = struct.unpack_from('<BBH', b, addr + 4)
        assert ntype == 0
        pos = addr + 8 + 16             # skip siblings
        pos += 8                        # key 0
        for _ in range(nused):
            child = struct.unpack_from('<Q', b, pos)[0]
            pos += 16                   # child + next key
            if level > 0:
                self._walk_group_btree(child, heap, links)
            else:
                self._read_snod(child, heap, links)

    def _read_snod(self, addr, heap, links):
        b = self.buf
        assert b[addr:addr + 4] == b'SNOD'
        nsym = struct.unpack_from('<H', b, addr + 6)[0]
        pos = addr + 8
        for _ in range(nsym):
            name_off, hdr = struct.unpack_from('<QQ', b, pos)
            pos += 40
            end = heap.index(b'\x00', name_off)
            links[heap[name_off:end].decode()] = hdr

    # ------------------------------------------------------------------ messages
    @staticmethod
    def _parse_dataspace(data):
        ver, rank, flags = struct.unpack_from('<BBB', data, 0)
        if ver == 1:
            pos = 8
        elif ver == 2:
            pos = 4
            if data[3] == 2:            # null dataspace
                return None
        else:
            raise ValueError('dataspace version %d' % ver)
        return tuple(struct.unpack_from('<%dQ' % rank, data, pos)) if rank else ()

    def _parse_datatype(self, data):
        cls = data[0] & 0x0F
        bits0 = data[1]
        size = struct.unpack_from('<I', data, 4)[0]
        order = '>' if (bits0 & 1) else '<'
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return ('num', np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size)))
        if cls == 1:
            return ('num', np.dtype('%sf%d' % (order, size)))
        if cls == 3:
            return ('num', np.dtype('S%d' % size))
        if cls == 9:
            vtype = bits0 & 0x0F        # 0 sequence, 1 string
            base = self._parse_datatype(data[8:])
            return ('vlen_str' if vtype == 1 else 'vlen', base, size)
        raise ValueError('unsupported datatype class %d' % cls)

    def _global_heap_object(self, addr, index):
        b = self.buf
        assert b[addr:addr + 4] == b'GCOL', 'bad global heap'
        csize = struct.unpack_from('<Q', b, addr + 8)[0]
        pos = addr + 16
        end = addr + csize
        while pos + 16 <= end:
            idx, _rc, _r, osize = struct.unpack_from('<HHIQ', b, pos)
            if idx == 0:
                break
            if idx == index:
                return b[pos + 16: pos + 16 + osize]
            pos += 16 + ((osize + 7) // 8) * 8
        raise KeyError('global heap object %d not found' % index)

    def _decode(self, dtype, shape, raw):
        kind = dtype[0]
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if kind == 'num':
            arr = np.frombuffer(raw, dtype=dtype[1], count=n).reshape(shape)
            return arr.astype(dtype[1].newbyteorder('='))
        if kind == 'vlen_str':
            out = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from('<IQI', raw, 16 * i)
                out.append(self._global_heap_object(gaddr, gidx)[:ln].decode())
            return out[0] if not shape else np.array(out, dtype=object).reshape(shape)
        raise ValueError('cannot decode %r' % (dtype,))

    def _attributes(self, msgs):
        attrs = {}
        for mtype, _f, data in msgs:
            if mtype != 0x0C:
                continue
            ver = data[0]
            nsz, tsz, ssz = struct.unpack_from('<HHH', data, 2)
            pos = 8
            if ver == 3:
                pos = 9
            pad = (lambda x: ((x + 7) // 8) * 8) if ver == 1 else (lambda x: x)
            name = data[pos:pos + nsz].split(b'\x00')[0].decode()
            pos += pad(nsz)
            dtype = self._parse_datatype(data[pos:pos + tsz])
            pos += pad(tsz)
            shape = self._parse_dataspace(data[pos:pos + ssz])
            pos += pad(ssz)
            attrs[name] = self._decode(dtype, shape, data[pos:])
        return attrs

    # ------------------------------------------------------------------ datasets
    def keys(self):
        return list(self.links.keys())

    def __contains__(self, name):
        return name in self.links

    def __getitem__(self, name):
        msgs = self._read_object_header(self.links[name])
        shape = dtype = layout = None
        filters = []
        for mtype, _f, data in msgs:
            if mtype == 0x01:
                shape = self._parse_dataspace(data)
            elif mtype == 0x03:
                dtype = self._parse_datatype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                filters = self._parse_filters(data)
        if dtype is None or layout is None:
            raise ValueError('%s is not a dataset' % name)
        if shape is None:
            return None
        if dtype[0] != 'num':
            raise ValueError('only numeric datasets supported')
        dt = dtype[1]
        n = int(np.prod(shape)) if shape else 1
        lver, lclass = layout[0], layout[1]
        if lver != 3:
            raise ValueError('layout version %d' % lver)
        if lclass == 0:                 # compact
            sz = struct.unpack_from('<H', layout, 2)[0]
            raw = layout[4:4 + sz]
        elif lclass == 1:               # contiguous
            addr, sz = struct.unpack_from('<QQ', layout, 2)
            raw = b'' if addr == _UNDEF else self.buf[addr:addr + sz]
            if addr == _UNDEF:
                return np.zeros(shape, dtype=dt.newbyteorder('='))
        elif lclass == 2:               # chunked
            rank = layout[2]
            btree = struct.unpack_from('<Q', layout, 3)[0]
            cdims = struct.unpack_from('<%dI' % rank, layout, 11)
            out = np.zeros(shape, dtype=dt.newbyteorder('='))
            if btree != _UNDEF and n:
                self._read_chunks(btree, rank, cdims[:-1], dt, filters, out)
            return out
        else:
            raise ValueError('layout class %d' % lclass)
        return self._decode(dtype, shape, raw)

    @staticmethod
    def _parse_filters(data):
        ver, nf = data[0], data[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid, nlen, _fl, ncv = struct.unpack_from('<HHHH', data, pos)
            pos += 8
            if ver == 1:
                pos += ((nlen + 7) // 8) * 8
            else:
                pos += nlen if fid >= 256 else 0
            cvals = struct.unpack_from('<%dI' % ncv, data, pos)
            pos += 4 * ncv
            if ver == 1 and ncv % 2:
                pos += 4
            out.append((fid, cvals))
        return out

    def _read_chunks(self, addr, rank, cdims, dt, filters, out):
        b = self.buf
        assert b[addr:addr + 4] == b'TREE', 'bad chunk btree'
        ntype, level, nused = struct.unpack_from('<BBH', b, addr + 4)
        assert ntype == 1
        keysz = 8 + 8 * rank
        pos = addr + 24
        for _ in range(nused):
            csize, fmask = struct.unpack_from('<II', b, pos)
            offs = struct.unpack_from('<%dQ' % rank, b, pos + 8)
            child = struct.unpack_from('<Q', b, pos + keysz)[0]
            pos += keysz + 8
            if level > 0:
                self._read_chunks(child, rank, cdims, dt, filters, out)
                continue
            raw = b[child:child + csize]
            for i, (fid, cvals) in reversed(list(enumerate(filters))):
                if fmask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    esz = cvals[0] if cvals else dt.itemsize
                    arr = np.frombuffer(raw, dtype=np.uint8)
                    raw = arr.reshape(esz, -1).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise ValueError('unsupported filter %d' % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs[:-1], cdims, out.shape):
                hi = min(o + c, s)
                sl_out.append(slice(o, hi))
                sl_in.append(slice(0, hi - o))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]


def load(path):
    """Return `(datasets: dict name -> ndarray, attrs: dict)` for a flat HDF5 file."""
    f = H5File(path)
    return {k: f[k] for k in f.keys()}, f.attrs
