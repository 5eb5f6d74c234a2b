"""Mirror of cont/PNGFile.java:19-658 (PNG / APNG chunk model around zlib streams).

IDAT chunks are concatenated into one zlib stream, every fdAT frame into its own, and zTXt / iCCP /
iTXt payloads are separate zlib streams.  On write every stream is re-emitted as a single chunk at the
position of the first chunk it came from and APNG sequence numbers are renumbered.
"""
import zlib

from ._io import ByteReader
from .deflate_files_container import DeflateFilesContainer
from .zlib_file import ZLibFile

PNG_SIG = bytes([137, 80, 78, 71, 13, 10, 26, 10])
INT_MAX = 2147483647


def _strlen(data, offset):  # Util.strlen (base/util/Util.java)
    i = offset
    n = len(data)
    while i < n and data[i] != 0:
        i += 1
    return i - offset


class PNGChunk:
    def __init__(self, ctype=b"\0\0\0\0", data=b""):
        self.type = bytes(ctype)
        self.data = bytearray(data)
        self.seqNum = 0

    def isIDAT(self): return self.type == b"IDAT"
    def isIEND(self): return self.type == b"IEND"
    def iszTXt(self): return self.type == b"zTXt"
    def isiCCP(self): return self.type == b"iCCP"
    def isiTXt(self): return self.type == b"iTXt"
    def isacTL(self): return self.type == b"acTL"
    def isfcTL(self): return self.type == b"fcTL"
    def isfdAT(self): return self.type == b"fdAT"
    def hasSeq(self): return self.isfdAT() or self.isfcTL()
    def isZLibCompressedNonIdat(self): return self.iszTXt() or self.isiCCP() or self.isiTXt()

    def setSeq(self, seq):  # PNGFile.java:67-73
        self.seqNum = seq
        self.data[0:4] = (seq & 0xffffffff).to_bytes(4, "big")

    def getZLibCompressedNonIdat(self):  # PNGFile.java:83-115
        if not self.isZLibCompressedNonIdat():
            return None
        data = self.data
        offset = _strlen(data, 0) + 2
        if self.isiTXt():
            if data[offset - 1] != 1:
                return None
            offset += 1
        if data[offset - 1] != 0:
            print("Only deflate compression is currently supported for %s chunks (read method %d)"
                  % (self.type.decode("latin-1"), data[offset - 1]))
            return None
        if self.isiTXt():
            offset += _strlen(data, offset) + 1
            offset += _strlen(data, offset) + 1
        return bytes(data[offset:])

    def setZLibCompressedNonIdat(self, new_zlib):  # PNGFile.java:117-138
        if not self.isZLibCompressedNonIdat():
            return
        data = self.data
        offset = _strlen(data, 0) + 2
        if self.isiTXt():
            offset += 1
            offset += _strlen(data, offset) + 1
            offset += _strlen(data, offset) + 1
        self.data = bytearray(data[:offset]) + bytearray(new_zlib)

    def write(self):  # PNGFile.java:140-158
        crc = zlib.crc32(bytes(self.data), zlib.crc32(self.type)) & 0xffffffff
        return len(self.data).to_bytes(4, "big") + self.type + bytes(self.data) + crc.to_bytes(4, "big")

    def read(self, r):  # PNGFile.java:162-215
        rd = lambda: r.read() & 0xff
        length = (rd() << 24) + (rd() << 16) + (rd() << 8) + rd()
        if length > INT_MAX:
            return False
        self.type = bytes([rd(), rd(), rd(), rd()])
        if length > 0:
            got = r.read_n(length)
            # Util.readFromInputStream pads with (byte) -1 past EOF
            self.data = bytearray(got) + bytearray(b"\xff" * (length - len(got)))
        else:
            self.data = bytearray()
        if self.hasSeq():
            self.seqNum = int.from_bytes(self.data[0:4], "big")
        crc = (rd() << 24) + (rd() << 16) + (rd() << 8) + rd()
        calc = zlib.crc32(bytes(self.data), zlib.crc32(self.type)) & 0xffffffff
        return calc == crc


class _PNGChunkHelper:  # PNGFile.java:413-572
    def __init__(self, stream_cls):
        self.stream_cls = stream_cls
        self.helperIdat = None
        self.helperFdats = None
        self.helperNonIDAT = []  # insertion-ordered (chunk, container) pairs (LinkedHashMap)
        self.outOfOrder = False
        self.readingfdAT = False
        self.readingIDAT = False
        self.seenacTL = False
        self.seenIDAT = False
        self.seenIEND = False
        self.baos = bytearray()
        self.seq = 0
        self.callFailed = False

    def shouldFlush(self, chunk):
        return (self.readingIDAT and not chunk.isIDAT()) or (self.readingfdAT and (chunk.isfcTL() or chunk.isIEND()))

    def flush(self):
        if self.readingIDAT:
            self.helperIdat = ZLibFile(self.stream_cls)
            if not self.helperIdat.read(bytes(self.baos)):
                return False
            self.helperIdat.deflateStream.setName("IDAT chunk")
            self.seenIDAT = True
            self.readingIDAT = False
        else:
            if self.helperFdats is None:
                self.helperFdats = []
            fdat = ZLibFile(self.stream_cls)
            self.helperFdats.append(fdat)
            if not fdat.read(bytes(self.baos)):
                return False
            fdat.deflateStream.setName("fdAT chunk %d" % len(self.helperFdats))
            self.readingfdAT = False
        self.baos = bytearray()
        return True

    def setReadState(self, chunk):
        if chunk.isfdAT():
            if not self.seenIDAT or self.readingIDAT:
                return False
            self.readingfdAT = True
        elif chunk.isIDAT():
            if self.seenIDAT or self.readingfdAT:
                return False
            self.readingIDAT = True
        return True

    def submitChunkImpl(self, chunk):
        if self.seenIEND:
            self.outOfOrder = True
            return False
        if chunk.hasSeq():
            if (self.seenIDAT and not self.seenacTL) or chunk.seqNum != self.seq:
                self.outOfOrder = True
                return False
            self.seq += 1
        if self.shouldFlush(chunk) and not self.flush():
            return False
        if chunk.isIEND():
            self.seenIEND = True
            return True
        if chunk.isacTL():
            if self.seenIDAT or self.seenacTL:
                return False
            self.seenacTL = True
            return True
        if not self.setReadState(chunk):
            return False
        if len(chunk.data) > 0:
            if self.readingIDAT:
                self.baos += chunk.data
            elif self.readingfdAT and chunk.isfdAT():
                self.baos += chunk.data[4:]
            elif chunk.isZLibCompressedNonIdat():
                z = chunk.getZLibCompressedNonIdat()
                if z is not None:
                    cont = ZLibFile(self.stream_cls)
                    if cont.read(z):
                        cont.deflateStream.setName(chunk.type.decode("latin-1") + " chunk")
                        self.helperNonIDAT.append((chunk, cont))
        return True

    def submitChunk(self, chunk):
        if self.callFailed:
            return False
        if not self.submitChunkImpl(chunk):
            self.callFailed = True
            return False
        return True

    def goodEndState(self):
        return (not self.callFailed and self.seenIEND and self.seenIDAT and not self.readingIDAT
                and not self.readingfdAT and not self.outOfOrder and (self.helperFdats is None or self.seenacTL))


class PNGFile(DeflateFilesContainer):
    def __init__(self, stream_cls=None):
        super().__init__(stream_cls)
        self.pngChunks = []
        self.idat = None
        self.fdats = None
        self.nonIDAT = []

    def getDeflateStreams(self):  # PNGFile.java:376-389
        out = list(self.idat.getDeflateStreams())
        if self.fdats is not None:
            for f in self.fdats:
                out += f.getDeflateStreams()
        for _, c in self.nonIDAT:
            out += c.getDeflateStreams()
        return out

    def read(self, data):  # PNGFile.java:574-605
        r = data if isinstance(data, ByteReader) else ByteReader(data)
        if r.read_n(8) != PNG_SIG:
            return False
        helper = _PNGChunkHelper(self.stream_cls)
        self.pngChunks = []
        ok = True
        while True:
            chunk = PNGChunk()
            if not chunk.read(r) or not helper.submitChunk(chunk):
                ok = False
                break
            self.pngChunks.append(chunk)
            if chunk.isIEND():
                break
        self.idat = helper.helperIdat
        self.fdats = helper.helperFdats
        self.nonIDAT = helper.helperNonIDAT
        return ok and helper.goodEndState()

    def _syncStreams(self):  # PNGFile.java:262-369
        first = next((c for c in self.pngChunks if c.isIDAT()), None)
        if first is None:
            raise IOError("No IDAT chunk found in PNG")
        index = self.pngChunks.index(first)
        self.pngChunks = [c for c in self.pngChunks if not c.isIDAT()]
        idat_bytes = self.idat.write()
        pos = 0
        while True:
            n = min(len(idat_bytes) - pos, INT_MAX)
            self.pngChunks.insert(index, PNGChunk(b"IDAT", idat_bytes[pos:pos + n]))
            index += 1
            pos += n
            if pos >= len(idat_bytes):
                break
        if self.fdats is not None:
            chunks = self.pngChunks
            i = 0  # ListIterator cursor
            for fdat in self.fdats:
                fdat_index = -1
                while i < len(chunks):
                    chunk = chunks[i]
                    i += 1
                    if fdat_index == -1:
                        if chunk.isfdAT():
                            fdat_index = i - 1
                            del chunks[i - 1]
                            i -= 1
                            fbytes = fdat.write()
                            fpos = 0
                            while True:
                                n = min(len(fbytes) - fpos, INT_MAX - 4)
                                chunks.insert(i, PNGChunk(b"fdAT", b"\0\0\0\0" + fbytes[fpos:fpos + n]))
                                i += 1
                                fpos += n
                                if fpos >= len(fbytes):
                                    break
                    else:
                        if chunk.isfdAT():
                            del chunks[i - 1]
                            i -= 1
                        if chunk.isfcTL():
                            break
                if fdat_index == -1:
                    raise IOError("Incorrect chunk order in APNG")
            seq = 0
            for chunk in chunks:
                if chunk.hasSeq():
                    chunk.setSeq(seq)
                    seq += 1
        for chunk, cont in self.nonIDAT:
            chunk.setZLibCompressedNonIdat(cont.write())

    def write(self):  # PNGFile.java:391-411
        self._syncStreams()
        out = bytearray(PNG_SIG)
        for c in self.pngChunks:
            out += c.write()
        return bytes(out)

    def fileType(self):
        return "PNG"


def _png_call(datas, merge_blocks, lib):
    from ._front import front_call
    return front_call("deft4cu_png_optimise_batch", datas, merge_blocks, lib)


def optimise_png_files(datas, merge_blocks=True, lib=None):
    """A LIST of PNG / APNG files through the native front-end (`deft4cu_png_optimise_batch`, csrc/png_front.cpp): the
    chunk model of this module done by host threads in C++, the zlib streams of all files as one device batch.
    Returns one dict per file (see `_front.front_optimise`): status, out, saved_bits, streams."""
    from ._front import front_optimise
    return front_optimise("deft4cu_png_optimise_batch", datas, merge_blocks, lib)
