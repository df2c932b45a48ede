"""gravity/sources.py:4-21 mirror."""


class SourceField:
    def __init__(self, receiver_list=None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.receiver_list = [] if receiver_list is None else list(receiver_list)
