"""ORACLE (test infrastructure). The handful of batchgenerators==0.21 file helpers the reference's
loss / module files import at module scope (reference nnunet_ext/paths.py:6)."""
import json
import os
import pickle

join = os.path.join
isdir = os.path.isdir
isfile = os.path.isfile


def maybe_mkdir_p(d):
    os.makedirs(d, exist_ok=True)


def load_pickle(f, mode="rb"):
    with open(f, mode) as fh:
        return pickle.load(fh)


def write_pickle(o, f, mode="wb"):
    with open(f, mode) as fh:
        pickle.dump(o, fh)


save_pickle = write_pickle


def load_json(f):
    with open(f) as fh:
        return json.load(fh)


def save_json(o, f, indent=4, sort_keys=True):
    with open(f, "w") as fh:
        json.dump(o, fh, indent=indent, sort_keys=sort_keys)
