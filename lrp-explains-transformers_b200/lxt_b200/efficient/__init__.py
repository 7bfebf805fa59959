"""mirror of lxt.efficient (filled in below)"""
