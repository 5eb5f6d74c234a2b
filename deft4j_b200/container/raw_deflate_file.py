"""Mirror of cont/RawDeflateFile.java:11-36."""
from ._io import ByteReader
from .deflate_files_container import DeflateFilesContainer


class RawDeflateFile(DeflateFilesContainer):
    def __init__(self, stream_cls=None):
        super().__init__(stream_cls)
        self.deflateStream = None

    def getDeflateStreams(self):
        return [self.deflateStream]

    def read(self, data):
        r = data if isinstance(data, ByteReader) else ByteReader(data)
        self.deflateStream = self.stream_cls()
        return self.deflateStream.parse(r)

    def write(self):
        out = self.deflateStream.asBytes()
        return out

    def fileType(self):
        return "Raw deflate stream"
