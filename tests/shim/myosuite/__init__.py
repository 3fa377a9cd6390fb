"""TEST INFRASTRUCTURE: stand-in for MyoSuite 1.2.3 (third party, not in the container), restated from memory; physics is
the fp64 oracle. See ../README.md."""
__version__ = "1.2.3-shim"
