"""`python -m wot_b200 optimal_transport ...` -- the one wot sub-command on the hot path (wot/__main__.py:7-23)."""
import sys


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    commands = ["optimal_transport"]
    if not argv or argv[0] in ("-h", "--help") or argv[0] not in commands:
        print("usage: wot_b200 {%s} [options]" % ",".join(commands))
        return 1 if argv and argv[0] not in ("-h", "--help") else 0
    from .commands import optimal_transport as cmd
    cmd.main(cmd.create_parser().parse_args(argv[1:]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
