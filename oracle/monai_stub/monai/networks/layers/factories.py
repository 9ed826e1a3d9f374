class Act:
    PRELU = "PRELU"
    LEAKYRELU = "LEAKYRELU"


class Norm:
    INSTANCE = "INSTANCE"
    BATCH = "BATCH"
