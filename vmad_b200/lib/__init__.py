"""Operator libraries (mirror of vmad/lib): fastpm, linalg, mpi."""
