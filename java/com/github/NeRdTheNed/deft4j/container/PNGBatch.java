// Panama FFM binding of deft4cu_png_optimise_batch (deft4cu_zip_optimise_batch has the same signature and result record:
// bind it the same way for ZipFile, ZipFile.java:46-127): PNGFile.read / optimise / write (deft4j-container's PNGFile.java:574-605,
// :262-369, :391-411 and DeflateFilesContainer.java:18-43) for a LIST of files in one native call.
// NOT compiled or tested in this repository's image (no JDK): see INTEGRATION.md.  JDK 22+.
package com.github.NeRdTheNed.deft4j.container;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.nio.charset.StandardCharsets;
import java.util.ArrayList;
import java.util.List;
import static java.lang.foreign.ValueLayout.*;

public final class PNGBatch {
    /** One file's outcome: what CMDUtil.optimiseFile prints and writes for it (cmd/CMDUtil.java:57-181). */
    public static final class Result {
        public int status;                 // 0 ok, 1 PNGFile.read returned false, 2 write() threw, 3 internal limit
        public long savedBits;             // DeflateFilesContainer.optimise's return value
        public byte[] out;                 // PNGFile.write()
        public String[] streamName;        // "IDAT chunk", "fdAT chunk 2", "zTXt chunk": getDeflateStreams() order
        public long[] streamSaved;
    }

    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("deft4cu.lib", "libdeft4cu.so"), Arena.global());
    private static MethodHandle h(String n, FunctionDescriptor d) { return L.downcallHandle(LIB.find(n).orElseThrow(), d); }
    private static final MethodHandle RUN  = h("deft4cu_png_optimise_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle FREE = h("deft4cu_free_file_results", FunctionDescriptor.ofVoid(ADDRESS, JAVA_INT));

    // struct deft4cu_file_result (include/deft4cu.h)
    private static final StructLayout RES = MemoryLayout.structLayout(
        JAVA_INT.withName("status"), JAVA_INT.withName("n_streams"), JAVA_LONG.withName("saved_bits"),
        ADDRESS.withName("out"), JAVA_LONG.withName("out_len"), ADDRESS.withName("stream_saved"), ADDRESS.withName("stream_name"));

    private PNGBatch() { }

    public static List<Result> optimise(List<byte[]> files, boolean mergeBlocks) throws Throwable {
        final int n = files.size();
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ptrs = a.allocate(ADDRESS, Math.max(1, n)), lens = a.allocate(JAVA_LONG, Math.max(1, n));
            for (int i = 0; i < n; i++) {
                byte[] f = files.get(i);
                MemorySegment m = a.allocate(Math.max(1, f.length));
                m.copyFrom(MemorySegment.ofArray(f));
                ptrs.setAtIndex(ADDRESS, i, m);
                lens.setAtIndex(JAVA_LONG, i, f.length);
            }
            MemorySegment res = a.allocate(RES, Math.max(1, n));
            int rc = (int) RUN.invokeExact(ptrs, lens, n, mergeBlocks ? 1 : 0, res);
            if (rc != 0) throw new IllegalStateException("deft4cu_png_optimise_batch failed: " + rc);
            List<Result> out = new ArrayList<>(n);
            try {
                for (int i = 0; i < n; i++) {
                    MemorySegment r = res.asSlice(i * RES.byteSize(), RES.byteSize());
                    Result x = new Result();
                    x.status = r.get(JAVA_INT, 0);
                    int ns = r.get(JAVA_INT, 4);
                    x.savedBits = r.get(JAVA_LONG, 8);
                    if (x.status == 0) {
                        long len = r.get(JAVA_LONG, 24);
                        x.out = r.get(ADDRESS, 16).reinterpret(len).toArray(JAVA_BYTE);
                        MemorySegment saved = r.get(ADDRESS, 32).reinterpret(8L * ns), names = r.get(ADDRESS, 40).reinterpret(8L * ns);
                        x.streamSaved = new long[ns];
                        x.streamName = new String[ns];
                        for (int k = 0; k < ns; k++) {
                            x.streamSaved[k] = saved.getAtIndex(JAVA_LONG, k);
                            x.streamName[k] = names.getAtIndex(ADDRESS, k).reinterpret(65536).getString(0, StandardCharsets.ISO_8859_1);
                        }
                    }
                    out.add(x);
                }
            } finally {
                FREE.invokeExact(res, n);
            }
            return out;
        }
    }
}
