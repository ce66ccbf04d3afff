from pmesh.pm import SelfComm

COMM_SELF = SelfComm()
COMM_WORLD = SelfComm()
SUM = 'sum'
