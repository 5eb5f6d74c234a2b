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
