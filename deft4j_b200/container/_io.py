"""Minimal java.io.InputStream stand-in used by the container mirrors."""


class ByteReader:
    """Sequential reader over a bytes object; read() returns -1 at EOF like InputStream.read()."""

    def __init__(self, data, pos=0):
        self.data = bytes(data)
        self.pos = pos

    def read(self):
        if self.pos >= len(self.data):
            return -1
        b = self.data[self.pos]
        self.pos += 1
        return b

    def read_n(self, n):
        out = self.data[self.pos:self.pos + n]
        self.pos += len(out)
        return out

    def remaining(self):
        return self.data[self.pos:]
