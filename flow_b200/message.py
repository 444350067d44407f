"""`with Message('...'):` -- indented log scopes (flow/message.py:12-24)."""
from .dolfin import begin, end


class Message(object):
    def __init__(self, string, verbose=True):
        self.string = string
        self.verbose = verbose

    def __enter__(self):
        if self.verbose:
            begin(self.string)

    def __exit__(self, tpe, value, traceback):
        if self.verbose:
            end()
