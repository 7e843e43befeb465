"""Minimal omegaconf stand-in (yaml.safe_load) so the reference's config plumbing imports."""
import yaml
from . import listconfig  # noqa: F401


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return yaml.safe_load(f)

    @staticmethod
    def create(obj):
        return obj
