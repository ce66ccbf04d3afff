"""Error types of the tape VM.

The names are API (user code and the reference's tests catch them): the same hierarchy as
vmad/core/error.py:1-21 -- everything derives from ``ModelError``; ``ExecutionError`` wraps an
exception raised inside a node together with the node that raised it.
"""


class ModelError(Exception):
    """base of every error raised while building or running a model"""


def _declare(*names):
    for name in names:
        globals()[name] = type(name, (ModelError,), {'__module__': __name__})


# argument parsing / graph construction / execution problems
_declare('BadArgument', 'UnpackError', 'DuplicatedOutput', 'MissingArgument',
         'OverwritePrecaution', 'UnexpectedOutput', 'ResolveError', 'InferError',
         'BrokenPrimitive')


class ExecutionError(ModelError):
    def __init__(self, msg, reason):
        ModelError.__init__(self, msg)
        self.reason = reason


def makeExecutionError(msg, reason):
    """an error that is an ExecutionError AND an instance of the original exception's type, so
    that ``except TypeError`` style handlers keep working when errors are wrapped"""
    try:
        both = type("ExecutionError(%s)" % type(reason).__name__, (type(reason), ExecutionError), {})
        return both(msg, reason)
    except TypeError:  # exception types with incompatible layouts
        return ExecutionError(msg, reason)
