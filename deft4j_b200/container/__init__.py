"""Host-side container wrappers (mirror of deft4j-container)."""
from .deflate_files_container import DeflateFilesContainer, optimise_streams, optimise_containers, read_containers
from .gz_file import GZFile, optimise_gz_files
from .png_file import PNGFile, optimise_png_files
from .raw_deflate_file import RawDeflateFile
from .zlib_file import ZLibFile, optimise_zlib_files
from .zip_file import ZipFile, optimise_zip_files
from .container_util import getContainerForExt, getContainerForBytes

__all__ = ["DeflateFilesContainer", "optimise_streams", "optimise_containers", "read_containers", "GZFile", "PNGFile", "optimise_png_files", "optimise_zip_files", "optimise_gz_files", "optimise_zlib_files", "RawDeflateFile", "ZLibFile", "ZipFile",
           "getContainerForExt", "getContainerForBytes"]
