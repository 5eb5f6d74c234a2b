"""Mirror of cont/ZLibFile.java:14-109 (zlib wrapper: CMF/FLG + one deflate stream + Adler-32)."""
import sys

from ._io import ByteReader
from .deflate_files_container import DeflateFilesContainer, RECALC


class ZLibFile(DeflateFilesContainer):
    def __init__(self, stream_cls=None):
        super().__init__(stream_cls)
        self.deflateStream = None
        self.CMF = 0
        self.FLG = 0
        self.adler32 = 0  # big-endian value as stored in the file

    def getDeflateStreams(self):
        return [self.deflateStream]

    def read(self, data):  # ZLibFile.java:59-95
        r = data if isinstance(data, ByteReader) else ByteReader(data)
        self.CMF = r.read()
        if (self.CMF & 0xF) != 8:
            print("ZLib non-deflate compression method %d not supported" % (self.CMF & 0xF), file=sys.stderr)
            return False
        self.FLG = r.read()
        if ((self.CMF << 8) + self.FLG) % 31 != 0:
            print("ZLib header check failed (FCHECK)", file=sys.stderr)
            return False
        if (self.FLG & 0x20) == 0x20:
            print("ZLib preset dictionary currently not supported", file=sys.stderr)
            return False
        self.deflateStream = self.stream_cls()
        if not self.deflateStream.parse(r):
            return False
        rd = lambda: r.read() & 0xff
        self.adler32 = (rd() << 24) + (rd() << 16) + (rd() << 8) + rd()
        return True

    def write(self):  # ZLibFile.java:33-57
        out = bytearray([self.CMF & 0xff, self.FLG & 0xff])
        out += self.deflateStream.asBytes()
        if RECALC:
            real = self.deflateStream.getChecksums()[1]
            if real != self.adler32:
                print("Warning: calculated Alder32 %d did not match expected Alder32 %d" % (real, self.adler32),
                      file=sys.stderr)
            out += (real & 0xffffffff).to_bytes(4, "big")
        else:
            out += (self.adler32 & 0xffffffff).to_bytes(4, "big")
        return bytes(out)

    def fileType(self):
        return "ZLib"

    def asGZipFiles(self):  # ZLibFile.java:102-107
        from .gz_file import GZFile
        gz = GZFile(self.stream_cls)
        gz.setData(self.deflateStream)
        return [gz]


def optimise_zlib_files(datas, merge_blocks=True, lib=None):
    """A LIST of zlib files through the native front-end (`deft4cu_zlib_optimise_batch`, csrc/gz_front.cpp)."""
    from ._front import front_optimise
    return front_optimise("deft4cu_zlib_optimise_batch", datas, merge_blocks, lib)
