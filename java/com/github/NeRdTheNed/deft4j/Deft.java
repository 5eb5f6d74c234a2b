// Panama FFM facade over libdeft4cu.so for deft4j-base's Deft (base/Deft.java:16-54).
// NOT compiled or tested in this repository's image (no JDK): see INTEGRATION.md.  JDK 22+.
package com.github.NeRdTheNed.deft4j;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

public final class Deft {
    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("deft4cu.lib", "libdeft4cu.so"), Arena.global());
    private static MethodHandle h(String n, FunctionDescriptor d) { return L.downcallHandle(LIB.find(n).orElseThrow(), d); }

    private static final MethodHandle OPT  = h("deft4cu_optimise_deflate_stream", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle FREE = h("deft4cu_free_buffer", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle SIZE = h("deft4cu_size_bits_fallback", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG));

    private Deft() { }

    /** Deft.optimiseDeflateStream(byte[]) (Deft.java:16-19) */
    public static byte[] optimiseDeflateStream(byte[] original) { return optimiseDeflateStream(original, true); }

    /** Deft.optimiseDeflateStream(byte[], boolean) (Deft.java:21-34): the SAME array comes back when nothing was saved,
     *  the stream does not parse, or anything goes wrong. */
    public static byte[] optimiseDeflateStream(byte[] original, boolean mergeBlocks) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocate(Math.max(1, original.length));
            in.copyFrom(MemorySegment.ofArray(original));
            MemorySegment out = a.allocate(ADDRESS), outLen = a.allocate(JAVA_LONG);
            int rc = (int) OPT.invokeExact(in, (long) original.length, mergeBlocks ? 1 : 0, out, outLen);
            MemorySegment p = out.get(ADDRESS, 0);
            if (rc != 0 || p.equals(MemorySegment.NULL)) return original;
            long n = outLen.get(JAVA_LONG, 0);
            byte[] result = p.reinterpret(n).toArray(JAVA_BYTE);
            FREE.invokeExact(p);
            return result;
        } catch (Throwable t) {
            return original;
        }
    }

    /** Deft.getSizeBitsFallback(byte[]) (Deft.java:48-54) */
    public static long getSizeBitsFallback(byte[] deflateStream) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = a.allocate(Math.max(1, deflateStream.length));
            in.copyFrom(MemorySegment.ofArray(deflateStream));
            return (long) SIZE.invokeExact(in, (long) deflateStream.length);
        } catch (Throwable t) {
            return deflateStream.length * 8L;
        }
    }
}
