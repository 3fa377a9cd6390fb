"""TEST INFRASTRUCTURE: makes the reference's env files importable here (see README.md in this directory)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "envs"))


def install():
    """Put the stand-in packages and the reference's ``src`` directory on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("the reference sources are not present: fixtures can only be regenerated where /root/reference exists")
    for p in (REPO, REFERENCE_SRC, HERE):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


def resolve_model_path(model_path: str) -> str:
    """The reference registers models under <ROOT_DIR>/data/myosuite/assets/...; the hand / elbow binaries are absent from
    the reference mount (.MISSING_LARGE_BLOBS) - those resolve to this repository's authored stand-ins of the same name."""
    if os.path.exists(model_path):
        return model_path
    parts = model_path.replace("\\", "/").split("/")
    alt = os.path.join(REPO, "myochallenge_b200", "assets", *parts[-2:])
    if os.path.exists(alt):
        return alt
    raise FileNotFoundError(model_path)
