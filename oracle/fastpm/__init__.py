"""Oracle stand-in for the ``fastpm`` python package (only ``fastpm.background``)."""
