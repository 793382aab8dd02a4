"""Run logger with the reference's contract (util/log.py:6-28): messages go to `save/{comment}/debug.log`
and to stderr in the format `[LEVEL|file:line] time > message`, the ROOT logger is set to DEBUG, and the
caller gets `(logger, log_dir, comment)` back.  Written for this package (not a copy): handlers are
de-duplicated so that train and valid can run in one process, and the root directory is a parameter.
"""
import logging
import pathlib
import time

LINE_FORMAT = '[%(levelname)s|%(filename)s:%(lineno)s] %(asctime)s > %(message)s'   # the log-line contract


def _default_comment():
    return time.strftime('%b%d_%H-%M-%S')      # same naming as the reference when no comment is given


def get_logger(comment=None, root='save'):
    comment = comment if comment is not None else _default_comment()
    directory = pathlib.Path(root) / comment
    directory.mkdir(parents=True, exist_ok=True)
    target = str((directory / 'debug.log').resolve())

    logger = logging.getLogger()
    logger.setLevel(logging.DEBUG)
    formatter = logging.Formatter(LINE_FORMAT)
    have_file = any(isinstance(h, logging.FileHandler) and getattr(h, 'baseFilename', None) == target
                    for h in logger.handlers)
    have_stream = any(type(h) is logging.StreamHandler and getattr(h, '_b200', False) for h in logger.handlers)
    if not have_file:
        handler = logging.FileHandler(target)
        handler.setFormatter(formatter)
        logger.addHandler(handler)
    if not have_stream:
        handler = logging.StreamHandler()
        handler.setFormatter(formatter)
        handler._b200 = True
        logger.addHandler(handler)
    return logger, '{0}/{1}'.format(root, comment), comment
