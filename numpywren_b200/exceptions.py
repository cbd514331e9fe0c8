"""Exception types of the LambdaPACK front end (same names as reference numpywren/exceptions.py)."""


class LambdaPackParsingException(Exception):
    pass


class LambdaPackTypeException(Exception):
    pass


class LambdaPackBackendGenerationException(Exception):
    pass


class LambdaPackRuntimeException(Exception):
    pass
