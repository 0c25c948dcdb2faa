"""Optional dependencies (mirror of the reference's src/mhealth/util/deps.py:1-9): when pandas is missing a stand-in
class keeps ``functools.singledispatch`` registrations importable."""
try:
    import pandas as pd
except ImportError:                                    # pragma: no cover
    class pd:                                          # noqa: N801
        class DataFrame(dict):
            pass

        class Series(list):
            pass
