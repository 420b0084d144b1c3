"""`df3d-cli` over the CUDA-backed Core: the reference's command line (df3d/cli.py:62-166) flag for flag, and its
run logic (cli.py:15-37, 170-326) -- single folder, --recursive (sub-folders named images/), --from-file, per-folder
error isolation, --skip-pose-estimation resume, --delete-images.

    python -m deepfly3d_b200.cli <INPUT> [--order 0 1 2 3 4 5 6] [-n N] [--skip-pose-estimation] ...

Video rendering (--video-2d / --video-3d) is visualisation and out of scope of this hot-path port (SURVEY.md
section 2, row 8): the flags are accepted so that existing command lines keep parsing, and an error names the
reference tool to use on the result pickle.  Two flags are added: --weights / --mean (the reference reads these
paths from its config, df3d/config.py:30-39).
"""
import argparse
import logging
from collections import deque
from pathlib import Path

logger = logging.getLogger("df3d.logger")


def parse_cli_args(argv=None):
    p = argparse.ArgumentParser(description="DeepFly3D pose estimation")
    p.add_argument("-v", "--verbose", help="Enable info output (such as progress bars)", action="store_true")
    p.add_argument("-vv", "--verbose2", help="Enable debug output", action="store_true")
    p.add_argument("-d", "--debug", help="Displays the argument list for debugging purposes", action="store_true")
    p.add_argument("input_folder", help="Without additional arguments, a folder containing unlabeled images.", metavar="INPUT")
    p.add_argument("--output-folder", default=None,
                   help="The name of the folder where results will be written. If not specified, a folder with the same name as "
                        "INPUT suffixed with '_df3d' will be created.")
    p.add_argument("-r", "--recursive", help="INPUT is a folder. Successively use its subfolders named 'images/'", action="store_true")
    p.add_argument("-f", "--from-file", action="store_true",
                   help="INPUT is a text-file, where each line names a folder. Successively use the listed folders.")
    p.add_argument("-x", "--delete-images", action="store_true",
                   help="Delete image files *after running df3d-cli*. Only deletes if there is corresponding .mp4 file is already in the folder.")
    p.add_argument("-n", "--num-images-max", default=0, type=int,
                   help="Maximal number of images to process. If 0 or not defined, process all images.")
    p.add_argument("--order", "--camera-ids", default=[0, 1, 2, 3, 4, 5, 6], type=int, nargs="*",
                   help="Ordering of the cameras provided as a list of ids. Example: --order 0 1 4 3 2 5 6.")
    p.add_argument("--video-2d", help="Generate pose2d videos", action="store_true")
    p.add_argument("--video-3d", help="Generate pose3d videos", action="store_true")
    p.add_argument("--skip-pose-estimation", help="Skip 2D and 3D pose estimation", dest="skip_estimation", action="store_true")
    p.add_argument("--batch-size", type=int, default=8,
                   help="Batch size for inference - how many images are processed through the model at once")
    p.add_argument("--pin-memory-disabled", action="store_true", help="Whether to disable `pin_memory` in the loader.")
    p.add_argument("--output-fps", type=float, default=None, help="FPS for output videos.")
    p.add_argument("--weights", default=None, help="hourglass checkpoint (sh8_deepfly.tar layout); default $DF3D_B200_WEIGHTS")
    p.add_argument("--mean", default=None, help="per-channel mean or the path of a mean.pth.tar; default $DF3D_B200_MEAN or 0.5")
    args = p.parse_args(argv)
    args.input_folder = Path(args.input_folder).expanduser().resolve()
    if args.output_folder is None:
        args.output_folder = args.input_folder.with_name(args.input_folder.stem + "_df3d")
    else:
        args.output_folder = Path(args.output_folder).expanduser().resolve()
    args.input_folder, args.output_folder = str(args.input_folder), str(args.output_folder)
    return args


def setup_logger(args):
    handler = logging.StreamHandler()
    handler.setLevel(logging.DEBUG)
    logger.addHandler(handler)
    logger.setLevel(logging.DEBUG if args.verbose2 else logging.INFO if args.verbose else logging.WARNING)


def print_debug(args):
    print(f"Enabled logging level: {logging.getLevelName(logger.getEffectiveLevel())}")
    print("Arguments are:")
    for key, val in vars(args).items():
        print(f"\t{key}: {val}")
    print()
    return 0


def find_subfolders(path, name):
    """Breadth-first search for sub-folders called `name`; a match is not descended into (cli.py:329-354)."""
    found, to_visit, visited = [], deque([Path(path)]), set()
    while to_visit:
        cur = to_visit.popleft()
        if cur.is_dir() and cur not in visited:
            visited.add(cur)
            if cur.name == name:
                found.append(str(cur))
            else:
                to_visit.extend(cur.iterdir())
    return found


def run(args):
    """One folder (cli.py:276-326)."""
    from .core import Core

    if args.skip_estimation and not args.video_2d and not args.video_3d:
        logger.info("Nothing to do. Check your command-line arguments.")
        return 0
    if args.video_2d or args.video_3d:
        raise NotImplementedError("--video-2d / --video-3d render with matplotlib in the reference (df3d/video.py) and are out of "
                                  "scope here: run the reference's df3d-cli --skip-pose-estimation --video-* on the result pickle")
    logger.info(f"\nWorking in {args.input_folder}")
    core = Core(args.input_folder, args.output_folder, args.num_images_max, args.order, weights=args.weights, mean=args.mean)
    if not args.skip_estimation:
        core.pose2d_estimation(args.batch_size, args.pin_memory_disabled)
        core.save()
    core.calibrate_calc(0, core.max_img_id)
    core.save()
    if args.delete_images:
        core.delete_images()
    return 0


def run_in_folders(args, folders):
    """Every folder in turn; an exception in one is logged and the others still run (cli.py:244-273)."""
    errors = []
    for folder in folders:
        try:
            args.input_folder = str(folder)
            args.output_folder = str(Path(folder).with_name(Path(folder).stem + "_df3d")) if args.auto_output else args.output_folder
            run(args)
        except KeyboardInterrupt:
            logger.warning("Keyboard Interrupt received. Terminating...")
            break
        except Exception as e:
            errors.append((folder, e))
            logger.error(f"An error occured while processing {folder}. Continuing...")
    if errors:
        logger.error(f"\n{len(errors)} out of {len(folders)} folders terminated with errors.")
        for folder, exc in errors:
            logger.error(f"\nIn {folder}", exc_info=exc)
    return 1 if errors else 0


def run_from_file(args):
    try:
        with open(args.input_folder, "r") as f:
            folders = [line.strip() for line in f]
    except FileNotFoundError:
        logger.error(f"Unable to find the file {args.input_folder}")
        return 1
    except IsADirectoryError:
        logger.error(f"{args.input_folder} is a directory, please provide a file instead.")
        return 1
    folders = [Path(f) for f in dict.fromkeys(folders) if f.strip()]          # unique, non-blank, in order
    bad = [f for f in folders if not f.is_dir()]
    for f in bad:
        logger.error(f"[Error] Not a directory or does not exist: {f}")
    if bad:
        return 1
    args.from_file = False
    return run_in_folders(args, folders)


def run_recursive(args):
    subfolders = find_subfolders(args.input_folder, "images")
    logger.info(f"Found {len(subfolders)} subfolder(s):\n-" + "\n-".join(subfolders))
    args.recursive = False
    return run_in_folders(args, subfolders)


def main(argv=None):
    import sys

    args = parse_cli_args(argv)
    args.auto_output = "--output-folder" not in (argv if argv is not None else sys.argv)
    setup_logger(args)
    if args.debug:
        return print_debug(args)
    if args.from_file and args.recursive:
        logger.error('Error: choose an input method between "from file" and "recursive" but not both.')
        return 1
    if args.recursive:
        return run_recursive(args)
    if args.from_file:
        return run_from_file(args)
    return run(args)


if __name__ == "__main__":
    raise SystemExit(main())
