"""Minimal stand-in for timm==0.5.4 (the version the reference pins: InvPT/README.md:57,
TaskPrompter/README.md:75), only the pieces the reference's hot path imports. TEST INFRASTRUCTURE:
used by oracle/ref_loader.py to import the unmodified reference models in this container."""
__version__ = "0.5.4-shim"
