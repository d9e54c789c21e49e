"""Mirror of ``PyContact.cy_modules`` (the native neighbour search), GPU-backed."""
