"""Mirror of cont/ContainerUtil.java:48-175: magic sniffing and extension → container mapping.

ZIP: the reference delegates parsing to the un-vendored lljzip dependency; zip_file.py restates the standard
layout and RecalculatingZipWriter (parity unpinned).
"""
from .gz_file import GZFile
from .png_file import PNGFile
from .raw_deflate_file import RawDeflateFile
from .zlib_file import ZLibFile
from .zip_file import ZipFile

_MAGICS = [
    (bytes([0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A]), "png"),
    (bytes([0x50, 0x4B, 0x03, 0x04]), "zip"),
    (bytes([0x50, 0x4B, 0x05, 0x06]), "zip"),
    (bytes([0x50, 0x4B, 0x07, 0x08]), "zip"),
    (bytes([0x1F, 0x8B]), "gzip"),
    (bytes([0x78, 0x01]), "zlib"),
    (bytes([0x78, 0x5E]), "zlib"),
    (bytes([0x78, 0x9C]), "zlib"),
    (bytes([0x78, 0xDA]), "zlib"),
]

_GZ_EXT = {"gz", "gzip", "tgz", "taz", "svgz", "cpgz", "wmz", "emz", "dat", "nbt", "mine", "mclevel"}
_ZLIB_EXT = {"zlib", "zz"}
_ZIP_EXT = {"zip", "jar", "apk", "ipa", "ear", "war", "epub"}


def detectFormat(data):
    """ContainerUtil.detectFormat (:63-86): first magic (in table order per byte) fully matched."""
    possible = [True] * len(_MAGICS)
    longest = max(len(m) for m, _ in _MAGICS)
    for n in range(longest):
        b = data[n] if n < len(data) else 0xff
        for i, (magic, name) in enumerate(_MAGICS):
            if possible[i]:
                if len(magic) < n + 1 or b != magic[n]:
                    possible[i] = False
                elif len(magic) <= n + 1:
                    return name
    return None


def getContainerForExt(ext, stream_cls=None):
    """ContainerUtil.getContainerForExt (:139-175)."""
    e = ext.lower().lstrip(".")
    if e in _GZ_EXT:
        return GZFile(stream_cls)
    if e in _ZLIB_EXT:
        return ZLibFile(stream_cls)
    if e in _ZIP_EXT:
        return ZipFile(stream_cls)
    if e == "png":
        return PNGFile(stream_cls)
    return RawDeflateFile(stream_cls)


def getContainerForBytes(data, filename="", stream_cls=None):
    """ContainerUtil.getContainerForPath (:89-131) minus the OS content-type probe."""
    fmt = detectFormat(data)
    if fmt is not None:
        return getContainerForExt(fmt, stream_cls)
    ext = filename.rsplit(".", 1)[1] if "." in filename else ""
    return getContainerForExt(ext, stream_cls)
