"""Mirror of cont/GZFile.java:13-226 (gzip member wrapper around one deflate stream)."""
import sys
import zlib

from ._io import ByteReader
from .deflate_files_container import DeflateFilesContainer, RECALC

FTEXT, FHCRC, FEXTRA, FNAME, FCOMMENT = 1, 2, 4, 8, 16


def _read_cstr(r):
    """Util.readStr (base/util/Util.java:232-241): bytes up to the NUL (platform charset round trip)."""
    out = bytearray()
    while True:
        b = r.read()
        if b == 0:
            break
        if b < 0:  # the reference would loop writing -1 forever; treat EOF as terminator
            break
        out.append(b)
    return bytes(out)


class GZFile(DeflateFilesContainer):
    def __init__(self, stream_cls=None):
        super().__init__(stream_cls)
        self.compressionMethod = 8
        self.flags = 0
        self.time = 0
        self.extraFlags = 0
        self.os = 255
        self.extra = None
        self.filename = None
        self.comment = None
        self.crc16 = 0
        self.deflateStream = None
        self.crc32 = 0
        self.isize = 0

    def read(self, data):  # GZFile.java:42-87
        r = data if isinstance(data, ByteReader) else ByteReader(data)
        if r.read() != 0x1f or r.read() != 0x8b:
            return False
        self.compressionMethod = r.read()
        if self.compressionMethod != 8:
            return False
        self.flags = r.read()
        if self.flags & 0xe0:
            return False
        self.time = r.read() + (r.read() << 8) + (r.read() << 16) + (r.read() << 24)
        self.extraFlags = r.read()
        self.os = r.read()
        if self.flags & FEXTRA:
            xlen = r.read() + (r.read() << 8)
            self.extra = r.read_n(xlen)
        if self.flags & FNAME:
            self.setFilename(_read_cstr(r))
        if self.flags & FCOMMENT:
            self.comment = _read_cstr(r)
        if self.flags & FHCRC:
            self.crc16 = r.read() + (r.read() << 8)
        name = self.filename.decode("latin-1") if self.filename else None
        self.deflateStream = self.stream_cls(name)
        if not self.deflateStream.parse(r):
            return False
        rd = lambda: r.read() & 0xff
        self.crc32 = rd() + (rd() << 8) + (rd() << 16) + (rd() << 24)
        self.isize = rd() + (rd() << 8) + (rd() << 16) + (rd() << 24)
        return True

    def write(self):  # GZFile.java:92-152
        out = bytearray([0x1f, 0x8b, self.compressionMethod & 0xff, self.flags & 0xff])
        out += (self.time & 0xffffffff).to_bytes(4, "little")
        out += bytes([self.extraFlags & 0xff, self.os & 0xff])
        if self.flags & FEXTRA:
            out += len(self.extra).to_bytes(2, "little") + self.extra
        if self.flags & FNAME:
            out += self.filename + b"\0"
        if self.flags & FCOMMENT:
            out += self.comment  # written WITHOUT its NUL terminator (GZFile.java:117-119)
        if self.flags & FHCRC:
            out += (self.crc16 & 0xffff).to_bytes(2, "little")
        body = self.deflateStream.asBytes()
        out += body
        if RECALC:
            real_crc, real_isize = self.deflateStream.getChecksums()[0:3:2]
            if real_crc != self.crc32:
                print("Warning: calculated CRC32 %d did not match expected CRC32 %d" % (real_crc, self.crc32),
                      file=sys.stderr)
            if self.isize != (real_isize & 0xffffffff):
                print("Warning: calculated size %d did not match expected size %d" % (real_isize, self.isize),
                      file=sys.stderr)
            out += (real_crc & 0xffffffff).to_bytes(4, "little")
            out += (real_isize & 0xffffffff).to_bytes(4, "little")
        else:
            out += (self.crc32 & 0xffffffff).to_bytes(4, "little")
            out += (self.isize & 0xffffffff).to_bytes(4, "little")
        return bytes(out)

    def setFilename(self, filename):  # GZFile.java:158-169
        self.filename = filename
        has = (self.flags & FNAME) != 0
        if filename:
            if not has:
                self.flags |= FNAME
        elif has:
            self.flags &= ~FNAME

    def setData(self, stream):  # GZFile.java:171-179
        self.deflateStream = stream
        self.compressionMethod = 8
        crc, _, size = stream.getChecksums()
        self.crc32 = crc
        self.isize = size & 0xffffffff

    def getDeflateStreams(self):
        return [self.deflateStream]

    def fileType(self):
        return "GZip"


def optimise_gz_files(datas, merge_blocks=True, lib=None):
    """A LIST of gzip files through the native front-end (`deft4cu_gz_optimise_batch`, csrc/gz_front.cpp): this module's
    header model in C++, the members' streams as one device batch, CRC-32 / ISIZE from the device.  Result records as
    `optimise_png_files`."""
    from ._front import front_optimise
    return front_optimise("deft4cu_gz_optimise_batch", datas, merge_blocks, lib)
