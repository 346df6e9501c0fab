"""sol::util for this path: asset lookup (src/util.rs:13-31 walks up to 5 parents for `assets/`)."""
import os


def find_asset(relative, start=None):
    d = os.path.abspath(start or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for _ in range(6):
        cand = os.path.join(d, "assets", relative)
        if os.path.exists(cand):
            return cand
        d = os.path.dirname(d)
    return None
