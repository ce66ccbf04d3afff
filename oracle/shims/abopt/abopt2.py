class VectorSpace(object):
    def __init__(self, addmul=None, dot=None):
        if addmul is not None:
            self.addmul = addmul
        if dot is not None:
            self.dot = dot


class Problem(object):
    def __init__(self, objective, gradient, hessian_vector_product=None, vs=None, **kwargs):
        self.objective = objective
        self.gradient = gradient
        self.hessian_vector_product = hessian_vector_product
        self.vs = vs
