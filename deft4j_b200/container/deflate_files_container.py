"""Mirror of deft4j-container's DeflateFilesContainer (cont/DeflateFilesContainer.java:14-100).

Containers stay host-side; they call the deflate stream model through the same
methods the reference uses: parse / optimise / write / getUncompressedData.  `stream_cls` selects the
stream implementation: the default is the CUDA-backed `deft4j_b200.DeflateStream`; the tests pass
the CPU oracle's stream class to pin the oracle against the reference's golden files.
"""
import sys

RECALC = True  # DeflateFilesContainer.java:15
PRINT_OPT = True  # Deft.java:11


def _default_stream_cls():
    from ..deflate_stream import DeflateStream
    return DeflateStream


def optimise_streams(streams, merge_blocks=True, out=sys.stdout):
    """static DeflateFilesContainer.optimise(List<DeflateStream>, boolean) (:18-43).

    Streams that share a batching backend (the CUDA DeflateStream) are optimised in one batched call;
    the printed lines and the return value are the reference's.
    """
    saved_total = 0
    if streams and hasattr(type(streams[0]), "optimise_batch"):
        saved_list = type(streams[0]).optimise_batch(streams, merge_blocks)
    else:
        saved_list = [s.optimise(merge_blocks) for s in streams]
    for i, (stream, saved) in enumerate(zip(streams, saved_list)):
        if PRINT_OPT and saved > 0 and out is not None:
            print("%d bits saved in stream %d (%s)" % (saved, i, stream.getName()), file=out)
        saved_total += saved
    if PRINT_OPT and saved_total > 0 and out is not None:
        print("Total bits saved %d" % saved_total, file=out)
    return saved_total


def read_containers(datas, names, stream_cls=None):
    """Read many files so that ALL their deflate streams are parsed in one device batch.

    A container parses its streams one at a time, because each parse tells it where the stream ended (the trailer
    follows).  Here every file is read twice on the host: a first pass with a recording stub collects the byte ranges
    the containers hand to `DeflateStream.parse`, those are parsed together (`DeflateStream.parse_batch`, one launch of
    every parse kernel), and the second pass hands the parsed streams back in order.  Returns one container per file,
    None where the file cannot be read.  Falls back to plain per-file reads for stream classes without `parse_batch`."""
    from .container_util import getContainerForBytes
    from ._io import ByteReader
    cls = stream_cls or _default_stream_cls()
    if not hasattr(cls, "parse_batch"):
        out = []
        for data, name in zip(datas, names):
            cont = getContainerForBytes(data, name, cls)
            out.append(cont if cont is not None and cont.read(data) else None)
        return out
    recorded = []

    class _Recorder:
        def __init__(self, name=None):
            self.name = name

        def parse(self, src):
            recorded[-1].append((self.name, src.remaining() if isinstance(src, ByteReader) else bytes(src)))
            return True

    for data, name in zip(datas, names):
        recorded.append([])
        try:
            cont = getContainerForBytes(data, name, _Recorder)
            if cont is None or not cont.read(data):
                recorded[-1] = None
        except Exception:  # noqa: BLE001 - a stub parse leaves the reader in front of the stream: later fields are garbage
            pass
    flat = [b for r in recorded if r for _, b in r]
    parsed = cls.parse_batch(flat, [n for r in recorded if r for n, _ in r]) if flat else []
    it = iter(parsed)
    out = []
    for data, name, rec in zip(datas, names, recorded):
        if rec is None:
            out.append(None)
            continue
        mine = [next(it) for _ in rec]
        queue = list(mine)

        def factory(nm=None, _q=queue):
            s = _q.pop(0) if _q else None
            return _Preparsed(s, nm, cls)

        cont = getContainerForBytes(data, name, factory)
        ok = cont is not None and cont.read(data)
        out.append(cont if ok else None)
    return out


class _Preparsed:
    """Stands in for a stream class while a container is read the second time: `parse` hands over the stream the
    batch parse produced (and advances the reader by what it consumed)."""

    def __new__(cls, stream, name, real_cls):
        from ._io import ByteReader
        if stream is None:
            stream = real_cls(name) if name is not None else real_cls()
            stream._preparse_failed = True
        elif name is not None:
            stream.setName(name)
        real_parse = stream.parse

        def parse(src, _s=stream):
            if getattr(_s, "_preparse_failed", False):
                return real_parse(src)   # (fails again, with the reference's behaviour)
            if isinstance(src, ByteReader):
                src.pos += _s.consumed
            return True

        stream.parse = parse
        return stream


def optimise_containers(containers, merge_blocks=True, outs=None):
    """Several containers (the files of `optimise-folder`, OptimiseFolder.java:33-67) as ONE stream list on the
    device; every container then reports exactly what its own `optimise` would have printed.  `outs`: one text sink
    per container (None: silent).  Returns the saved bits per container."""
    lists = [c.getDeflateStreams() for c in containers]
    flat = [s for l in lists for s in l]
    if flat and hasattr(type(flat[0]), "optimise_batch"):
        saved_flat = type(flat[0]).optimise_batch(flat, merge_blocks)
    else:
        saved_flat = [s.optimise(merge_blocks) for s in flat]
    totals, k = [], 0
    for ci, l in enumerate(lists):
        out = outs[ci] if outs is not None else None
        total = 0
        for i, stream in enumerate(l):
            saved = saved_flat[k]; k += 1
            if PRINT_OPT and saved > 0 and out is not None:
                print("%d bits saved in stream %d (%s)" % (saved, i, stream.getName()), file=out)
            total += saved
        if PRINT_OPT and total > 0 and out is not None:
            print("Total bits saved %d" % total, file=out)
        totals.append(total)
    return totals


class DeflateFilesContainer:
    def __init__(self, stream_cls=None):
        self.stream_cls = stream_cls or _default_stream_cls()

    def getDeflateStreams(self):
        raise NotImplementedError

    def read(self, data):
        """read(InputStream | byte[]) → bool"""
        raise NotImplementedError

    def write(self):
        """byte[] write() (:55-63): returns bytes, raises IOError when the container cannot be written."""
        raise NotImplementedError

    def optimise(self, merge_blocks=True, out=sys.stdout):
        return optimise_streams(self.getDeflateStreams(), merge_blocks, out)

    def getStreamInfo(self):
        lines = []
        streams = self.getDeflateStreams()
        for i, s in enumerate(streams):
            lines.append("Stream %d\n%s\n" % (i, s.printBlockInfo()))
        return ("File type: " + self.fileType() + "\nDeflate streams info:\n" + "".join(lines) +
                "Total streams: %d" % len(streams))

    def fileType(self):
        raise NotImplementedError
