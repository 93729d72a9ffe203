"""Strategy 'stream' (placeholder until the TMA streaming generator lands)."""


def applicable(program, domain, device_props):
    return False, "stream strategy not built yet"
