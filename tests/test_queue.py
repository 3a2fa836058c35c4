"""CPU-only: the block queue of the hand-off, scenario for scenario as reference test/test_queue.c."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

VP, SZ = C.c_void_p, C.c_size_t


@pytest.fixture()
def q(sdrm):
    lib = sdrm.lib
    lib.create_queue.argtypes = [C.c_uint32, C.c_uint16, C.c_bool, C.POINTER(VP)]
    lib.queue_put.argtypes = [VP, SZ, VP]
    lib.take_buffer_for_processing.argtypes = [C.POINTER(VP), C.POINTER(SZ), VP]
    lib.take_buffer_for_processing.restype = None
    lib.complete_buffer_processing.argtypes = [VP]
    lib.complete_buffer_processing.restype = None
    lib.interrupt_waiting_the_data.argtypes = [VP]
    lib.interrupt_waiting_the_data.restype = None
    lib.destroy_queue.argtypes = [VP]
    lib.destroy_queue.restype = None
    return lib


def make(lib, buffer_size, queue_size, blocking):
    h = VP()
    code = lib.create_queue(buffer_size, queue_size, blocking, C.byref(h))
    return code, h


def put(lib, h, values):
    a = np.array(values, dtype=np.float32)
    return lib.queue_put(a.ctypes.data_as(VP), len(a) // 2, h)


def take(lib, h):
    buf, n = VP(), SZ()
    lib.take_buffer_for_processing(C.byref(buf), C.byref(n), h)
    if not buf.value:
        return None
    out = np.frombuffer((C.c_char * (n.value * 8)).from_address(buf.value), dtype=np.float32).copy()
    lib.complete_buffer_processing(h)
    return out


def test_invalid_arguments(q):
    assert make(q, 4, 0, False)[0] == -1
    assert make(q, 0, 10, False)[0] == -1
    code, h = make(q, 4, 10, False)
    assert code == 0
    assert q.queue_put(None, 25, h) == -1
    assert put(q, h, []) == -1
    assert put(q, h, range(1, 11)) == -1  # 5 samples > buffer_size 4
    q.destroy_queue(h)


def test_terminated_only_after_fully_processed(q):
    code, h = make(q, 262144, 10, False)
    assert code == 0
    data = list(range(1, 11))
    assert put(q, h, data) == 0
    q.interrupt_waiting_the_data(h)
    assert np.array_equal(take(q, h), np.array(data, np.float32))
    assert take(q, h) is None
    q.interrupt_waiting_the_data(None)  # no-op
    q.destroy_queue(h)


def test_put_take(q):
    code, h = make(q, 262144, 10, False)
    assert put(q, h, range(1, 11)) == 0
    assert put(q, h, [1, 2]) == 0
    assert np.array_equal(take(q, h), np.arange(1, 11, dtype=np.float32))
    assert np.array_equal(take(q, h), np.array([1, 2], np.float32))
    q.destroy_queue(h)


def test_overflow_overwrites_the_newest_block(q):
    code, h = make(q, 262144, 1, False)
    assert put(q, h, range(1, 11)) == 0
    assert put(q, h, range(11, 21)) == 0
    assert np.array_equal(take(q, h), np.arange(11, 21, dtype=np.float32))
    q.destroy_queue(h)


def test_put_skipped_after_termination(q):
    code, h = make(q, 262144, 1, True)
    assert put(q, h, range(1, 11)) == 0
    q.interrupt_waiting_the_data(h)
    assert put(q, h, range(1, 11)) == -1
    q.destroy_queue(h)


def test_blocking_put_waits_for_a_free_buffer(q):
    code, h = make(q, 16, 2, True)
    assert put(q, h, [1, 2]) == 0 and put(q, h, [3, 4]) == 0
    done = []

    def producer():
        done.append(put(q, h, [5, 6]))

    t = threading.Thread(target=producer)
    t.start()
    time.sleep(0.2)
    assert not done  # still blocked: both buffers are filled
    assert np.array_equal(take(q, h), np.array([1, 2], np.float32))
    t.join(5)
    assert done == [0]
    assert np.array_equal(take(q, h), np.array([3, 4], np.float32))
    assert np.array_equal(take(q, h), np.array([5, 6], np.float32))
    q.destroy_queue(h)
