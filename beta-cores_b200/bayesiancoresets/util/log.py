"""Logging setup mirroring bayesiancoresets/util/log.py: the root logger is set to ERROR with a
stderr handler whose format carries the algorithm id (`extra={'id': ...}` through LoggerAdapter)."""
import logging
import sys

LOGLEVELS = dict(error=logging.ERROR, warning=logging.WARNING, critical=logging.CRITICAL, info=logging.INFO,
                 debug=logging.DEBUG, notset=logging.NOTSET)


def set_verbosity(verb):
    logging.getLogger().setLevel(LOGLEVELS[verb])


class _IdDefault(logging.Filter):
    def filter(self, record):           # records from other libraries carry no id: do not crash the formatter
        if not hasattr(record, 'id'):
            record.id = record.name
        return True


def _install():
    root = logging.getLogger()
    if any(getattr(h, '_betacores', False) for h in root.handlers):
        return
    h = logging.StreamHandler(stream=sys.stderr)
    h._betacores = True
    h.addFilter(_IdDefault())
    h.setFormatter(logging.Formatter('%(levelname)s - %(id)s.%(funcName)s(): %(message)s'))
    root.addHandler(h)
    root.setLevel(LOGLEVELS['error'])


_install()
