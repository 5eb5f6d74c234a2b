"""Mirror of cont/ZipFile.java:46-128 + cont/lljzip/RecalculatingZipWriter.java:23-136.

The reference parses archives with the third-party lljzip 2.3.0 (`ZipIO.readStandard`, un-vendored), so the reader
here restates the *standard* strategy from the ZIP application note: end-of-central-directory record found from the
back, central directory entries, each linked to the local header at its recorded offset.  Local files are kept in
offset order.  **Parity unpinned**: the reference ships no optimised ZIP golden; the written layout follows
RecalculatingZipWriter (which is in the reference): every local header is rewritten with CRC and sizes of its central
directory entry, data descriptors are dropped (the general purpose flag is written back unchanged), the central
directory follows with recomputed offsets, then the end record with recomputed counts, size and offset.
No Zip64, no multi-disk, no encryption: such archives are refused (`read` returns False), like the reference prints
"No local file headers detected ... data descriptors or Zip64".
"""
import struct
import sys

from .deflate_files_container import DeflateFilesContainer

SIG_LOCAL, SIG_CENTRAL, SIG_END = 0x04034B50, 0x02014B50, 0x06054B50
DEFLATED = 8


class _Local:
    __slots__ = ("version", "flags", "method", "mtime", "mdate", "crc32", "csize", "usize", "name", "extra", "data",
                 "offset", "central")


class _Central:
    __slots__ = ("made_by", "version", "flags", "method", "mtime", "mdate", "crc32", "csize", "usize", "name", "extra",
                 "comment", "disk", "iattr", "eattr", "offset", "local")


class ZipFile(DeflateFilesContainer):
    def __init__(self, stream_cls=None):
        super().__init__(stream_cls)
        self.locals = []
        self.centrals = []
        self.end_disk = 0
        self.end_start_disk = 0
        self.comment = b""
        self.streams = []          # (local header, DeflateStream) in local-file order (ZipFile.java:100-124)

    def fileType(self):
        return "Zip"

    def getDeflateStreams(self):
        return [s for _, s in self.streams]

    # ---- read (ZipFile.java:82-127) -----------------------------------------------------------------------
    def read(self, data):
        data = bytes(data.readall()) if hasattr(data, "readall") else bytes(data)
        end = data.rfind(struct.pack("<I", SIG_END))
        if end < 0 or end + 22 > len(data):
            print("No local file headers detected in zip file, may be due to use of data descriptors or Zip64 features",
                  file=sys.stderr)
            return False
        (_, disk, start_disk, n_here, n_total, cd_size, cd_off, clen) = struct.unpack_from("<IHHHHIIH", data, end)
        if n_total == 0xFFFF or cd_off == 0xFFFFFFFF or cd_size == 0xFFFFFFFF:
            print("No local file headers detected in zip file, may be due to use of data descriptors or Zip64 features",
                  file=sys.stderr)
            return False
        self.end_disk, self.end_start_disk = disk, start_disk
        self.comment = data[end + 22:end + 22 + clen]
        self.centrals, self.locals = [], []
        p = cd_off
        for _ in range(n_total):
            if p + 46 > len(data) or struct.unpack_from("<I", data, p)[0] != SIG_CENTRAL:
                break
            c = _Central()
            (_, c.made_by, c.version, c.flags, c.method, c.mtime, c.mdate, c.crc32, c.csize, c.usize, nlen, xlen, klen,
             c.disk, c.iattr, c.eattr, c.offset) = struct.unpack_from("<IHHHHHHIIIHHHHHII", data, p)
            c.name = data[p + 46:p + 46 + nlen]
            c.extra = data[p + 46 + nlen:p + 46 + nlen + xlen]
            c.comment = data[p + 46 + nlen + xlen:p + 46 + nlen + xlen + klen]
            c.local = None
            p += 46 + nlen + xlen + klen
            self.centrals.append(c)
        for c in self.centrals:
            o = c.offset
            if o + 30 > len(data) or struct.unpack_from("<I", data, o)[0] != SIG_LOCAL:
                continue
            l = _Local()
            (_, l.version, l.flags, l.method, l.mtime, l.mdate, l.crc32, l.csize, l.usize, nlen, xlen) = \
                struct.unpack_from("<IHHHHHIIIHH", data, o)
            l.name = data[o + 30:o + 30 + nlen]
            l.extra = data[o + 30 + nlen:o + 30 + nlen + xlen]
            l.offset, l.central = o, c
            # data descriptors leave the local sizes zero: take the central directory's (ZipFile.java:104-107)
            if l.csize == 0 and (l.csize, l.usize, l.crc32) != (c.csize, c.usize, c.crc32):
                l.csize, l.usize, l.crc32 = c.csize, c.usize, c.crc32
            start = o + 30 + nlen + xlen
            l.data = data[start:start + l.csize]
            c.local = l
            self.locals.append(l)
        self.locals.sort(key=lambda l: l.offset)
        if not self.locals:
            print("No local file headers detected in zip file, may be due to use of data descriptors or Zip64 features",
                  file=sys.stderr)
            return False
        self.streams = []
        for l in self.locals:
            if l.method != DEFLATED:          # stored / other methods are carried through untouched (:97-99)
                continue
            name = l.name.decode("utf-8", "replace") if l.name else None
            s = self.stream_cls(name) if name is not None else self.stream_cls()
            if not s.parse(l.data):
                print("Failed to parse stream for file %s" % name, file=sys.stderr)
                return False
            self.streams.append((l, s))
        return True

    # ---- write (ZipFile.java:46-79, RecalculatingZipWriter.java:23-136) --------------------------------------
    def write(self):
        for l, s in self.streams:             # syncStreams
            comp = s.asBytes()
            l.data = comp
            l.csize = len(comp)
            if l.central is not None:
                l.central.csize = len(comp)
        out = bytearray()
        new_off = {}
        for l in self.locals:
            c = l.central
            csize, usize, crc = (c.csize, c.usize, c.crc32) if c is not None else (l.csize, l.usize, l.crc32)
            new_off[c.offset if c is not None else l.offset] = len(out)
            out += struct.pack("<IHHHHHIIIHH", SIG_LOCAL, l.version, l.flags, l.method, l.mtime, l.mdate, crc,
                               csize & 0xFFFFFFFF, usize & 0xFFFFFFFF, len(l.name), len(l.extra))
            out += l.name + l.extra + l.data
        start_central = len(out)
        n = 0
        for c in self.centrals:
            if c.offset not in new_off:
                raise IOError("Error writing CentralDirectoryFileHeader %r, could not find old offset" % c.name)
            csize, usize = (c.local.csize, c.local.usize) if c.local is not None else (c.csize, c.usize)
            out += struct.pack("<IHHHHHHIIIHHHHHII", SIG_CENTRAL, c.made_by, c.version, c.flags, c.method, c.mtime,
                               c.mdate, c.crc32, csize & 0xFFFFFFFF, usize & 0xFFFFFFFF, len(c.name), len(c.extra),
                               len(c.comment), c.disk, c.iattr, c.eattr, new_off[c.offset])
            out += c.name + c.extra + c.comment
            n += 1
        central_size = len(out) - start_central
        out += struct.pack("<IHHHHIIH", SIG_END, self.end_disk, self.end_start_disk, n, n, central_size, start_central,
                           len(self.comment))
        out += self.comment
        return bytes(out)


def optimise_zip_files(datas, merge_blocks=True, lib=None):
    """A LIST of ZIP archives through the native front-end (`deft4cu_zip_optimise_batch`, csrc/zip_front.cpp): the archive
    model of this module in C++, the method-8 entries of all archives as one device batch.  Same result records as
    `optimise_png_files`; stream names are the entries' file names."""
    from ._front import front_optimise
    return front_optimise("deft4cu_zip_optimise_batch", datas, merge_blocks, lib, name_encoding="utf-8")
