"""mirror of lxt.explicit"""
