COMM_WORLD = None
SUM = None
