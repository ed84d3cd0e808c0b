"""A byte buffer whose last byte is followed by an inaccessible page: a parser that reads past the end of its input
dies with SIGSEGV instead of silently passing (test infrastructure)."""
import ctypes as C
import mmap

PAGE = mmap.PAGESIZE
_libc = C.CDLL(None, use_errno=True)
_libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]


class Guarded:
    def __init__(self, capacity):
        self.pages = (capacity + PAGE - 1) // PAGE + 1
        self.map = mmap.mmap(-1, (self.pages + 1) * PAGE)
        self.base = C.addressof(C.c_char.from_buffer(self.map))
        if _libc.mprotect(self.base + self.pages * PAGE, PAGE, 0) != 0:
            raise OSError(C.get_errno(), "mprotect")

    def put(self, data):
        """Copy data so that it ends exactly at the guard page; returns its address."""
        n = len(data)
        addr = self.base + self.pages * PAGE - n
        C.memmove(addr, bytes(data), n)
        return addr
