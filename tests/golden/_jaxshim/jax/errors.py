class TracerBoolConversionError(Exception):
    pass
