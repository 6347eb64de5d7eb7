"""Stand-in for variant-extractor ^4.0.6 (not installed).  TEST INFRASTRUCTURE ONLY."""
from .variants import VariantRecord, VariantType  # noqa: F401

_REGISTRY = {}


def register_vcf(path, records):
    _REGISTRY[path] = list(records)


class VariantExtractor:
    def __init__(self, path, *a, **kw):
        self._records = _REGISTRY[path]

    def __iter__(self):
        return iter(self._records)

    def close(self):
        pass
