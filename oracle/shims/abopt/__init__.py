"""Stand-in for abopt (absent): the two base-class names that
/root/reference/vmad/contrib/chisquare.py:17 imports.  TEST INFRASTRUCTURE ONLY."""
