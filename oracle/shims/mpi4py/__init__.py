"""Stand-in for mpi4py (absent from this image): a one-rank communicator only.
TEST INFRASTRUCTURE ONLY -- lets /root/reference's test modules import."""
