"""``ds_io`` entry point in the reference's argument dialect (reference ``nvme/test_ds_aio.py``)."""
from .ds_aio_args import get_validated_args
from .ds_aio_basic import aio_basic_multiprocessing
from .ds_aio_handle import aio_handle_multiprocessing


def ds_io_main(argv=None):
    args = get_validated_args(argv)
    print(f"Testing deepspeed_aio python frontend: {'read' if args.read else 'write'} {args.io_size} bytes, "
          f"block {args.block_size}, queue depth {args.queue_depth}, {args.multi_process} process(es)")
    fn = aio_handle_multiprocessing if (args.handle or args.gpu or args.use_gds or True) and not getattr(args, "basic", False) \
        else aio_basic_multiprocessing
    return fn(args, args.read)


def main():
    ds_io_main()


if __name__ == "__main__":
    main()
