"""An in-memory stand-in for the few h5py calls the reference's ``fit`` makes.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py to run the UNMODIFIED reference's
``LevelSetMachineLearning.fit`` in a container without h5py).  Files live in a process-wide
dictionary keyed by path; groups hold datasets (numpy arrays read / written with ``[...]``) and
an ``attrs`` dictionary.  Call sites: lsml/core/datasets_handler.py:86-202, 333, 370-482;
lsml/core/temporary_data_handler.py:100; lsml/core/fit_job_handler.py:210-521.
"""
import numpy

_FILES = {}


class _Dataset:
    def __init__(self, data):
        self._data = numpy.array(data)

    def __getitem__(self, key):
        return numpy.array(self._data[key])

    def __setitem__(self, key, value):
        self._data[key] = value

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype


class _Group:
    def __init__(self):
        self._items = {}
        self.attrs = {}

    def create_group(self, name):
        if name in self._items:
            raise ValueError("group %r already exists" % (name,))
        self._items[name] = _Group()
        return self._items[name]

    def create_dataset(self, name, data=None, **_ignored):       # compression=... is ignored
        self._items[name] = _Dataset(data)
        return self._items[name]

    def __getitem__(self, name):
        return self._items[name]

    def __contains__(self, name):
        return name in self._items

    def keys(self):
        return sorted(self._items.keys())                        # h5py lists keys in name order

    def __len__(self):
        return len(self._items)


class File(_Group):
    def __init__(self, name, mode="r"):
        name = str(name)
        if mode == "w" or (mode in ("a", "w-", "x") and name not in _FILES):
            _FILES[name] = _Group()
        elif name not in _FILES:
            raise OSError("Unable to open file %r (fake h5py: no such file)" % (name,))
        root = _FILES[name]
        self._items, self.attrs = root._items, root.attrs

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def reset():
    _FILES.clear()
