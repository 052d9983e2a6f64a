"""Config loading / logging helpers (drop-in for reference utils/file_io.py:8-72)."""
import importlib
import logging
import os
import sys

from segmentation3d.utils.attrdict import install_easydict_shim


def load_config(pyfile):
    """Import a python config file and return its `cfg` object (reference file_io.py:8-28)."""
    assert os.path.isfile(pyfile), 'The file {} does not exits!'.format(pyfile)
    install_easydict_shim()
    folder, base = os.path.split(os.path.abspath(pyfile))
    name = os.path.splitext(base)[0]
    sys.path.append(folder)
    try:
        sys.modules.pop(name, None)          # always re-read the file (the reference reloads it)
        module = importlib.import_module(name)
    finally:
        sys.path.pop()
    return module.cfg


def setup_logger(log_file, name):
    """Logger writing to stdout and `log_file` (reference file_io.py:31-58)."""
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    logger.handlers = []
    fmt = logging.Formatter('%(asctime)s %(message)s')
    for h in (logging.StreamHandler(sys.stdout), logging.FileHandler(log_file)):
        h.setFormatter(fmt)
        logger.addHandler(h)
    return logger


def readlines(path):
    """Non-empty stripped lines of a text file (reference file_io.py:61-72)."""
    with open(path, 'r') as f:
        return [ln.strip() for ln in f if ln.strip()]
