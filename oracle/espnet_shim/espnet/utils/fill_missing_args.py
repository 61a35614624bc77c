"""Stub of espnet.utils.fill_missing_args.fill_missing_args: attributes that
`add_arguments` defines and the Namespace lacks get the parser default."""
import argparse


def fill_missing_args(args, add_arguments):
    assert isinstance(args, argparse.Namespace) or args is None
    parser = argparse.ArgumentParser()
    add_arguments(parser)
    defaults, _ = parser.parse_known_args([])
    out = {} if args is None else dict(vars(args))
    for k, v in vars(defaults).items():
        out.setdefault(k, v)
    return argparse.Namespace(**out)
