// Panama FFM facade over libdeft4cu.so (include/deft4cu.h) with the public surface of deft4j-base's DeflateStream.
// NOT compiled or tested in this repository's image (no JDK): see INTEGRATION.md.  JDK 22+.
package com.github.NeRdTheNed.deft4j.deflate;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

/** Same public surface as the pure-Java DeflateStream (DeflateStream.java:18-182,492-660), backed by libdeft4cu.so. */
public final class DeflateStream implements AutoCloseable {
    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("deft4cu.lib", "libdeft4cu.so"), Arena.global());
    private static MethodHandle h(String n, FunctionDescriptor d) { return L.downcallHandle(LIB.find(n).orElseThrow(), d); }

    private static final MethodHandle PARSE    = h("deft4cu_stream_parse",  FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle OPTIMISE = h("deft4cu_stream_optimise", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle OPT_BATCH= h("deft4cu_stream_optimise_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle WRITE    = h("deft4cu_stream_write",  FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    private static final MethodHandle SIZE     = h("deft4cu_stream_size_bits", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    private static final MethodHandle ULEN     = h("deft4cu_stream_uncompressed_len", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    private static final MethodHandle UDATA    = h("deft4cu_stream_uncompressed", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle SUMS     = h("deft4cu_stream_checksums", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle FREE     = h("deft4cu_stream_free",   FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LASTERR  = h("deft4cu_last_error",    FunctionDescriptor.of(ADDRESS));

    private MemorySegment handle = MemorySegment.NULL;
    private long consumed;          // bytes parse() used; containers continue reading their trailer here
    private String name;

    public DeflateStream(String name) { this.name = name; }
    public String getName() { return name; }  public void setName(String n) { name = n; }
    public long getConsumedBytes() { return consumed; }

    /** DeflateStream.parse(byte[]) (:68).  The InputStream overload (:72) reads the remaining bytes, calls this, and
     *  pushes back data.length - consumed bytes (PushbackInputStream) so GZFile.read (:80-85) finds its trailer. */
    public boolean parse(byte[] data) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocate(Math.max(1, data.length)); in.copyFrom(MemorySegment.ofArray(data));
            MemorySegment out = a.allocate(ADDRESS), used = a.allocate(JAVA_LONG);
            int rc = (int) PARSE.invokeExact(in, (long) data.length, out, used);
            if (rc == 4 /* DEFT4CU_ERR_CUDA */) throw new IllegalStateException(lastError());
            if (rc != 0) return false;                       // parse failure (:105,109) or unsupported (H10)
            handle = out.get(ADDRESS, 0); consumed = used.get(JAVA_LONG, 0);
            return true;
        }
    }
    /** DeflateStream.optimise(boolean) (:496-566) */
    public long optimise(boolean mergeBlocks) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment saved = a.allocate(JAVA_LONG);
            int rc = (int) OPTIMISE.invokeExact(handle, mergeBlocks ? 1 : 0, saved);
            if (rc != 0 && rc != 3) throw new IllegalStateException(lastError());
            return saved.get(JAVA_LONG, 0);
        }
    }
    /** static DeflateFilesContainer.optimise(List, boolean) (DeflateFilesContainer.java:18-43): one device batch. */
    public static long[] optimiseAll(java.util.List<DeflateStream> streams, boolean mergeBlocks) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            int n = streams.size();
            MemorySegment hs = a.allocate(ADDRESS, n), saved = a.allocate(JAVA_LONG, n);
            for (int i = 0; i < n; i++) hs.setAtIndex(ADDRESS, i, streams.get(i).handle);
            int rc = (int) OPT_BATCH.invokeExact(hs, n, mergeBlocks ? 1 : 0, saved);
            if (rc != 0 && rc != 3) throw new IllegalStateException(lastError());
            return saved.toArray(JAVA_LONG);
        }
    }
    /** asBytes() (:652-660) / write(OutputStream) (:128-145) */
    public byte[] asBytes() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment len = a.allocate(JAVA_LONG);
            if ((int) WRITE.invokeExact(handle, MemorySegment.NULL, 0L, len) != 0) throw new java.io.IOException(lastError());
            long n = len.get(JAVA_LONG, 0);
            MemorySegment dst = a.allocate(Math.max(1, n));
            if ((int) WRITE.invokeExact(handle, dst, n, len) != 0) throw new java.io.IOException(lastError());
            return dst.asSlice(0, n).toArray(JAVA_BYTE);
        }
    }
    public boolean write(java.io.OutputStream os) throws Throwable { os.write(asBytes()); return true; }
    public long getSizeBits() throws Throwable { return (long) SIZE.invokeExact(handle); }                  // :171-182
    public long getUncompressedLength() throws Throwable { return (long) ULEN.invokeExact(handle); }
    /** {crc32, adler32} of getUncompressedData() without moving it to the host (GZFile.java:130-145, ZLibFile.java:42-51) */
    public int[] getChecksums() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment c = a.allocate(JAVA_INT), d = a.allocate(JAVA_INT);
            if ((int) SUMS.invokeExact(handle, c, d) != 0) throw new IllegalStateException(lastError());
            return new int[] { c.get(JAVA_INT, 0), d.get(JAVA_INT, 0) };
        }
    }
    public byte[] getUncompressedData() throws Throwable {                                                   // :159-169
        try (Arena a = Arena.ofConfined()) {
            long n = getUncompressedLength();
            MemorySegment dst = a.allocate(Math.max(1, n));
            if ((int) UDATA.invokeExact(handle, dst, n) != 0) throw new IllegalStateException(lastError());
            return dst.asSlice(0, n).toArray(JAVA_BYTE);
        }
    }
    @Override public void close() throws RuntimeException {
        try { if (!handle.equals(MemorySegment.NULL)) FREE.invokeExact(handle); } catch (Throwable t) { throw new RuntimeException(t); }
        handle = MemorySegment.NULL;
    }
    private static String lastError() throws Throwable { return ((MemorySegment) LASTERR.invokeExact()).reinterpret(4096).getString(0); }
}
