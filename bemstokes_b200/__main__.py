"""`python -m bemstokes_b200 [start_frame [end_frame]]` — the reference's executable (source/main.cc:5-71): reads
`parameters_3.prm` from the working directory, runs frames start..end (defaults 0..139) through BEMProblem.run and
prints EXECUTION OK; any exception is reported like main.cc:48-71 and gives exit status 1.  Options beyond the
reference: --prm FILE, --device N, --output-dir DIR.  The third positional argument of the reference (compose = 1,
the post-processing `composer`) is outside this package."""
import argparse
import sys


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="python -m bemstokes_b200", description=__doc__)
    ap.add_argument("start_frame", nargs="?", type=int, default=0)
    ap.add_argument("end_frame", nargs="?", type=int, default=139)
    ap.add_argument("--prm", default="parameters_3.prm")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--output-dir", default=".")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(sys.argv[1:] if argv is None else argv)
    try:
        from . import BEMProblem
        p = BEMProblem(device=args.device)
        p.parse_parameters(args.prm)
        p.output_dir = args.output_dir
        p.run(args.start_frame, args.end_frame)
        print("EXECUTION OK")
    except Exception as exc:  # same reporting as main.cc:48-60
        bar = "-" * 52
        sys.stderr.write("\n\n%s\nException on processing: \n%s\nAborting!\n%s\n" % (bar, exc, bar))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
