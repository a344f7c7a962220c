"""Command-line arguments of the ``ds_io`` benchmark (reference ``nvme/ds_aio_args.py``)."""
import argparse
import os

from .test_ds_aio_utils import refine_integer_value

MAPPING_DELIMITER = ":"


def refine_args(args):
    if args.io_size and isinstance(args.io_size, str):
        args.io_size = refine_integer_value(args.io_size)
    if args.block_size and isinstance(args.block_size, str):
        args.block_size = refine_integer_value(args.block_size)
    return args


def _get_mapping_dict(args):
    """``--folder_to_device_mapping /mnt/a:0 /mnt/b:1`` -> {device id: [folders]}; without it every process uses --folder."""
    if args.folder is not None:
        return {i: [args.folder] for i in range(args.multi_process)}
    out = {}
    for m in args.folder_to_device_mapping:
        folder, dev = m.rsplit(MAPPING_DELIMITER, 1)
        out.setdefault(int(dev), []).append(folder)
    return out


def _validate_folder_mapping(args):
    errs = []
    for m in args.folder_to_device_mapping or []:
        parts = m.rsplit(MAPPING_DELIMITER, 1)
        if len(parts) != 2 or not parts[1].isdigit():
            errs.append(f"Invalid mapping '{m}' (expected <folder>{MAPPING_DELIMITER}<device id>)")
        elif not os.path.isdir(parts[0]):
            errs.append(f"Folder {parts[0]} in mapping '{m}' does not exist")
    return errs


def validate_args(args):
    errs = []
    if args.folder is not None and args.folder_to_device_mapping:
        errs.append("--folder and --folder_to_device_mapping cannot be specified together.")
    if args.folder is None and not args.folder_to_device_mapping:
        errs.append("At least one of --folder or --folder_to_device_mapping must be specified.")
    if args.folder is not None and not os.path.isdir(args.folder):
        errs.append(f"Invalid folder in --folder: {args.folder}")
    errs += _validate_folder_mapping(args)
    if args.use_gds and not args.gpu:
        errs.append("--gpu must be set to transfer with --use_gds")
    for e in errs:
        print(f"Error: {e}")
    return not errs


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(description="DeepSpeed async I/O benchmark")
    p.add_argument("--folder", default=None, type=str, help="Folder to use for I/O.")
    p.add_argument("--folder_to_device_mapping", default=None, type=str, nargs="+",
                   help="Mapping of folder to (gpu) device id, (ignored for cpu accesses). Can be specified multiple times.")
    p.add_argument("--io_size", type=str, default="256M", help="Number of bytes to read or write (K/M/G suffixes).")
    p.add_argument("--read", action="store_true", help="Perform read I/O (default is write)")
    p.add_argument("--multi_process", type=int, default=1, help="Number of parallel processes doing I/O (default 1).")
    p.add_argument("--block_size", type=str, default="1M", help="I/O block size (K/M/G suffixes).")
    p.add_argument("--queue_depth", type=int, default=32, help="I/O queue depth (default 32).")
    p.add_argument("--single_submit", action="store_true", help="Submit I/O requests one at a time (default is batched)")
    p.add_argument("--sequential_requests", action="store_true", help="Wait for each request before issuing the next")
    p.add_argument("--validate", action="store_true", help="Perform validation of I/O transfer in library.")
    p.add_argument("--handle", action="store_true", help="Use AIO handle (always true here; kept for CLI compatibility).")
    p.add_argument("--loops", type=int, default=3, help="Count of operation repetitions")
    p.add_argument("--io_parallel", type=int, default=1, help="Per iop parallelism (threads per handle)")
    p.add_argument("--gpu", action="store_true", help="Use GPU memory")
    p.add_argument("--use_gds", action="store_true", help="Enable GPUDirect Storage")
    p.add_argument("--slow_bounce_buffer", action="store_true", help="For GPU memory transfers, measure impact of bounce buffer pinning")
    return p.parse_args(argv)


def get_validated_args(argv=None):
    args = refine_args(parse_arguments(argv))
    if not validate_args(args):
        raise SystemExit(1)
    args.mapping_dict = _get_mapping_dict(args)
    args.mapping_list = [(dev, f) for dev, folders in args.mapping_dict.items() for f in folders]
    return args
