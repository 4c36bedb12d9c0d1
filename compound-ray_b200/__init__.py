"""compound-ray_b200: B200-native drop-in for CompoundRay's libEyeRenderer3 (compound-eye render path).

The directory name carries a hyphen (it mirrors the reference project name), so import it through
``importlib`` -- see ``__graft_entry__.load_package()`` -- or put this directory on ``sys.path`` and
``import eye_renderer``.
"""
