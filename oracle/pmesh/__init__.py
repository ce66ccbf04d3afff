"""Oracle stand-in for the ``pmesh`` package (see ``oracle/__init__.py``)."""
