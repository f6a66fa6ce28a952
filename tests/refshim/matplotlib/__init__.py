"""refshim matplotlib: PNG output is out of scope; every pyplot call is a no-op."""
