"""Minimal HDF5 writer / reader for the NMFk result files (``results.h5``, data_io.py:199-209 of the reference) on hosts
without ``h5py``.

Writes the oldest, simplest on-disk structures of the HDF5 File Format Specification (the ones ``h5py`` itself produces
with ``libver='earliest'``): version-0 superblock, one old-style root group (version-1 object header with a symbol-table
message -> version-1 group B-tree -> symbol-table nodes -> local heap with the link names) and, per dataset, a version-1
object header with dataspace (v1), datatype (v1; IEEE little-endian float32 / float64 or signed integers), fill-value
(v2, default fill value) and contiguous data-layout (v3) messages followed by the raw little-endian array.  Flat files only
(datasets directly under ``/``, at most 256), no attributes, no chunking, no compression -- exactly what
``h5py.File(..., 'w').create_dataset(name, data=array)`` of the reference's writer needs.

``read`` parses the same subset (plus compact layout, the version-1/2 layout messages of libhdf5 <= 1.6 and a user block
in front of the superblock) and is what ``data_io.read_results`` uses when ``h5py`` is not importable.  The authoring
image has no h5py / libhdf5, so the pin is the one libhdf5-written file it does hold, scipy's
``io/matlab/tests/data/testhdf5_7.4_GLNX86.mat`` (MATLAB v7.3 = earliest-format HDF5): ``tests/test_h5min.py`` reads it
with ``read`` and checks that ``write`` produces byte-identical datatype / dataspace messages, B-tree node prefix and
heap layout for the same dataset; the fill-value (v2) and layout (v3) messages follow the specification text only.
When ``h5py`` is present ``data_io`` uses it instead of this module.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'
LEAF_K, INTERNAL_K = 4, 16                     # group B-tree fan-outs recorded in the superblock (library defaults)
SNOD_ENTRIES = 2 * LEAF_K                      # symbols per symbol-table node
HEAP_FREE_NULL = 1                             # "no next free block" marker of the local heap


def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack('<HHB3s', mtype, len(data), flags, b'\x00\x00\x00') + data


def _object_header(messages):
    body = b''.join(messages)
    # version 1 prefix: version, reserved, number of messages, reference count, header size, 4 bytes of alignment padding
    return struct.pack('<BBHII4s', 1, 0, len(messages), 1, len(body), b'\x00' * 4) + body


def _datatype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        expo_loc, expo_size, mant_size, bias = (23, 8, 23, 127) if dt.itemsize == 4 else (52, 11, 52, 1023)
        # class 1 (floating point), version 1; bit field: little-endian, mantissa normalisation 2 (msb implied), sign bit
        head = struct.pack('<BBBBI', 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        props = struct.pack('<HHBBBBI', 0, dt.itemsize * 8, expo_loc, expo_size, 0, mant_size, bias)
        return head + props
    if dt.kind in 'iu' and dt.itemsize in (1, 2, 4, 8):
        # class 0 (fixed point), version 1; bit 3 of the bit field = signed
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize)
        props = struct.pack('<HH', 0, dt.itemsize * 8)
        return head + props
    raise TypeError('h5min supports float32/float64 and integer datasets, not %s' % dt)


def _dataset_header(arr, data_address):
    shape = arr.shape
    dataspace = struct.pack('<BBBBI', 1, len(shape), 0, 0, 0) + b''.join(struct.pack('<Q', d) for d in shape)
    fill = struct.pack('<BBBBI', 2, 2, 2, 1, 0)        # v2: allocate late, write if set, defined with size 0 (= default)
    layout = struct.pack('<BBQQ', 3, 1, data_address, arr.nbytes)     # v3, class 1 = contiguous
    return _object_header([_msg(0x0001, dataspace), _msg(0x0003, _datatype_message(arr.dtype), flags=1),
                           _msg(0x0005, fill), _msg(0x0008, layout)])


def write(path, datasets):
    """``datasets``: {name: array-like}.  Scalars become 0-dimensional datasets, like ``create_dataset(name, data=x)``."""
    items = []
    for name, val in datasets.items():
        a = np.asarray(val)
        if a.dtype.kind == 'f' and a.dtype.itemsize not in (4, 8):
            a = a.astype(np.float64)
        if a.dtype.kind == 'O':                        # a list of numpy scalars of mixed width and the like
            a = a.astype(np.float64)
        if a.dtype.kind == 'b':
            a = a.astype(np.int8)
        a = a.astype(a.dtype.newbyteorder('<')).copy(order='C')    # (ascontiguousarray would make a 0-d array 1-d)
        items.append((name.encode('ascii'), a))
    items.sort(key=lambda t: t[0])                     # symbol-table entries are ordered by link name (strcmp)
    assert 0 < len(items) <= SNOD_ENTRIES * 2 * INTERNAL_K, 'h5min: 1..256 datasets'

    # ---- local heap data segment: the empty name of the root at offset 0, then the link names, then one free block
    heap = bytearray(b'\x00' * 8)
    name_off = {}
    for name, _ in items:
        name_off[name] = len(heap)
        heap += _pad8(name + b'\x00')
    free_off = len(heap)
    heap_size = len(heap) + 32
    heap += struct.pack('<QQ', HEAP_FREE_NULL, heap_size - free_off) + b'\x00' * 16

    # ---- file layout (every structure 8-byte aligned)
    root_header = _object_header([_msg(0x0011, struct.pack('<QQ', 0, 0))])       # addresses patched below
    pos = 96                                           # superblock (56) + root symbol-table entry (40)
    addr_root = pos; pos += len(root_header)
    addr_btree = pos; btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8; pos += btree_size
    addr_heap = pos; pos += 32
    addr_heap_data = pos; pos += len(heap)
    groups = [items[i:i + SNOD_ENTRIES] for i in range(0, len(items), SNOD_ENTRIES)]
    snod_size = 8 + SNOD_ENTRIES * 40
    addr_snod = []
    for _ in groups:
        addr_snod.append(pos); pos += snod_size
    obj_addr, data_addr = {}, {}
    for name, a in items:
        hdr_len = len(_dataset_header(a, 0))
        obj_addr[name] = pos; pos += hdr_len
        data_addr[name] = pos if a.nbytes else UNDEF
        pos += a.nbytes + (-a.nbytes % 8)
    eof = pos

    out = bytearray()
    out += SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    out += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    out += struct.pack('<QQII', 0, addr_root, 1, 0) + struct.pack('<QQ', addr_btree, addr_heap)     # root entry, cached
    assert len(out) == 96
    out += _object_header([_msg(0x0011, struct.pack('<QQ', addr_btree, addr_heap))])
    # group B-tree, one leaf level: key[0] = "" (offset 0), child[i] = symbol-table node i, key[i+1] = its largest name
    bt = bytearray(b'TREE' + struct.pack('<BBHQQ', 0, 0, len(groups), UNDEF, UNDEF))
    bt += struct.pack('<Q', 0)
    for g, a_s in zip(groups, addr_snod):
        bt += struct.pack('<QQ', a_s, name_off[g[-1][0]])
    bt += b'\x00' * (btree_size - len(bt))
    out += bt
    out += b'HEAP' + struct.pack('<B3sQQQ', 0, b'\x00' * 3, heap_size, free_off, addr_heap_data)
    out += heap
    for g in groups:
        sn = bytearray(b'SNOD' + struct.pack('<BBH', 1, 0, len(g)))
        for name, _ in g:
            sn += struct.pack('<QQII16s', name_off[name], obj_addr[name], 0, 0, b'\x00' * 16)
        sn += b'\x00' * (snod_size - len(sn))
        out += sn
    for name, a in items:
        assert len(out) == obj_addr[name]
        out += _dataset_header(a, data_addr[name])
        raw = a.tobytes()
        out += raw + b'\x00' * (-len(raw) % 8)
    assert len(out) == eof
    with open(path, 'wb') as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------------------------
# reader (the same subset, parsed from the specification independently of the writer's layout decisions)
# ---------------------------------------------------------------------------------------------------------------
def _read_messages(buf, addr, base=0):
    version, _, nmsg, _, size = struct.unpack_from('<BBHII', buf, addr)
    if version != 1:
        raise ValueError('h5min reads version-1 object headers only')
    pos, end, msgs = addr + 16, addr + 16 + size, []
    while pos < end and len(msgs) < nmsg:
        mtype, msize, _flags = struct.unpack_from('<HHB', buf, pos)
        data = bytes(buf[pos + 8:pos + 8 + msize])
        if mtype == 0x0010:                             # object header continuation: (address, length)
            caddr, clen = struct.unpack_from('<QQ', data, 0)
            msgs += _read_block(buf, caddr + base, clen, nmsg - len(msgs) - 1)
        else:
            msgs.append((mtype, data))
        pos += 8 + msize
    return msgs


def _read_block(buf, pos, length, limit):
    end, msgs = pos + length, []
    while pos + 8 <= end and len(msgs) < limit:
        mtype, msize, _flags = struct.unpack_from('<HHB', buf, pos)
        msgs.append((mtype, bytes(buf[pos + 8:pos + 8 + msize])))
        pos += 8 + msize
    return msgs


def _dtype_of(data):
    cv, b0, b1, _b2, size = struct.unpack_from('<BBBBI', data, 0)
    cls = cv & 0x0F
    order = '>' if (b0 & 1) else '<'
    if cls == 1:
        return np.dtype(order + 'f%d' % size)
    if cls == 0:
        return np.dtype(order + ('i' if (b0 & 0x08) else 'u') + '%d' % size)
    raise ValueError('h5min reads fixed-point and floating-point datasets only (class %d)' % cls)


def _symbols(buf, btree_addr, heap_data, base=0):
    """(name, object header address) of every link below a version-1 group B-tree node."""
    sig, ntype, level, used = struct.unpack_from('<4sBBH', buf, btree_addr)
    if sig != b'TREE' or ntype != 0:
        raise ValueError('not a group B-tree node')
    out, pos = [], btree_addr + 24 + 8                  # skip key[0]
    for _ in range(used):
        child, _key = struct.unpack_from('<QQ', buf, pos)
        child += base
        pos += 16
        if level > 0:
            out += _symbols(buf, child, heap_data, base)
            continue
        ssig, _ver, _res, nsym = struct.unpack_from('<4sBBH', buf, child)
        if ssig != b'SNOD':
            raise ValueError('not a symbol-table node')
        for i in range(nsym):
            noff, oaddr = struct.unpack_from('<QQ', buf, child + 8 + 40 * i)
            end = heap_data.index(b'\x00', noff)
            out.append((heap_data[noff:end].decode('ascii'), oaddr))
    return out


def read(path):
    """{name: ndarray} of the datasets directly under the root group."""
    buf = memoryview(open(path, 'rb').read())
    sb = 0                                               # the superblock sits at 0 or, behind a user block, at 512, 1024, ...
    while bytes(buf[sb:sb + 8]) != SIGNATURE:
        sb = 512 if sb == 0 else sb * 2
        if sb + 96 > len(buf):
            raise ValueError('not an HDF5 file')
    buf = buf[sb:]
    if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
        raise ValueError('h5min reads version-0 superblocks with 8-byte offsets / lengths only')
    base = struct.unpack_from('<Q', buf, 24)[0] - sb     # addresses are relative to the base address (= the user block size)
    root_header = struct.unpack_from('<Q', buf, 56 + 8)[0] + base
    stab = [d for t, d in _read_messages(buf, root_header, base) if t == 0x0011]
    if not stab:
        raise ValueError('root group is not an old-style (symbol table) group')
    btree_addr, heap_addr = struct.unpack_from('<QQ', stab[0], 0)
    hsig, _v, _r, hsize, _free, hdata = struct.unpack_from('<4sB3sQQQ', buf, heap_addr + base)
    if hsig != b'HEAP':
        raise ValueError('bad local heap')
    heap_data = bytes(buf[hdata + base:hdata + base + hsize])
    out = {}
    for name, oaddr in _symbols(buf, btree_addr + base, heap_data, base):
        msgs = dict((t, d) for t, d in _read_messages(buf, oaddr + base, base) if t in (0x0001, 0x0003, 0x0008))
        if len(msgs) < 3:
            continue                                     # a sub-group or something this subset does not cover
        sp = msgs[0x0001]
        if sp[0] == 1:
            rank, dims_at = sp[1], 8
        elif sp[0] == 2:
            rank, dims_at = sp[1], 4
        else:
            raise ValueError('dataspace message version %d' % sp[0])
        shape = struct.unpack_from('<%dQ' % rank, sp, dims_at) if rank else ()
        dt = _dtype_of(msgs[0x0003])
        lay = msgs[0x0008]
        n = int(np.prod(shape)) if rank else 1
        if lay[0] == 3 and lay[1] == 1:                  # contiguous
            daddr, dsize = struct.unpack_from('<QQ', lay, 2)
            raw = bytes(buf[daddr + base:daddr + base + n * dt.itemsize]) if daddr != UNDEF else b''
        elif lay[0] == 3 and lay[1] == 0:                # compact: size (2 bytes) then the data
            dsize = struct.unpack_from('<H', lay, 2)[0]
            raw = lay[4:4 + dsize]
        elif lay[0] in (1, 2) and lay[2] == 1:           # versions 1 / 2 (libhdf5 <= 1.6): rank + 1, class, 5 reserved, address
            daddr = struct.unpack_from('<Q', lay, 8)[0]
            raw = bytes(buf[daddr + base:daddr + base + n * dt.itemsize]) if daddr != UNDEF else b''
        else:
            raise ValueError('h5min reads contiguous / compact layouts only')
        arr = np.frombuffer(raw, dtype=dt, count=n if raw else 0)
        out[name] = arr.reshape(shape).astype(dt.newbyteorder('='), copy=True) if raw else np.zeros(shape, dt.newbyteorder('='))
    return out
