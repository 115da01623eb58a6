"""Mirror of the hot-path parts of the reference's `volsdf.utils` (rend_util, general.get_class)."""
