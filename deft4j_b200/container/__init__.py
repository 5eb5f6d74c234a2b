"""Host-side container wrappers (mirror of deft4j-container)."""
from .deflate_files_container import DeflateFilesContainer, optimise_streams
from .gz_file import GZFile
from .png_file import PNGFile
from .raw_deflate_file import RawDeflateFile
from .zlib_file import ZLibFile
from .zip_file import ZipFile
from .container_util import getContainerForExt, getContainerForBytes

__all__ = ["DeflateFilesContainer", "optimise_streams", "GZFile", "PNGFile", "RawDeflateFile", "ZLibFile", "ZipFile",
           "getContainerForExt", "getContainerForBytes"]
