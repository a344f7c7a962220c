"""Is the native async-I/O engine usable on this machine? (reference ``nvme/validate_async_io.py``)"""


def main():
    from deepspeed_b200.op_builder import ALL_OPS
    b = ALL_OPS.get("async_io")
    ok = False
    try:
        bb = b() if isinstance(b, type) else b
        ok = bool(bb is not None and bb.is_compatible())
        if ok:
            from deepspeed_b200.ops.aio import aio_handle
            aio_handle()  # loads libdsb200_cpu.so and opens an engine
    except Exception as e:
        print(f"async_io check failed: {e}")
        ok = False
    print(f"DeepSpeed async_io: {'[OKAY]' if ok else '[FAIL]'}")
    assert ok, "the async I/O engine is not usable here"
    return ok


if __name__ == "__main__":
    main()
