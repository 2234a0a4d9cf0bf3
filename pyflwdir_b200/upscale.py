"""Placeholder for the reference's `pyflwdir.upscale` (out of scope of the D8 hot path). Importing it works,
so that code written against the reference package layout loads; using anything in it says where to go instead."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise NotImplementedError(f"upscale.{name} is outside the D8 hot path that pyflwdir_b200 accelerates; use Deltares/pyflwdir for it")
