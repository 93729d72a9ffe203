"""Semantic model of a stencil kernel (what the Python front end produces).

Keeps the node vocabulary of reference ``stencil_code/stencil_model.py:17-68``
(``InteriorPointsLoop``, ``NeighborPointsLoop``, ``MathFunction``, ``GridElement``) and adds the
handful of plain expression/statement nodes the reference borrowed from ctree
(``ctree.c.nodes``: Constant, SymbolRef, BinaryOp, ..., un-vendored there) so the model is
self-contained.  Nodes are inert records; all meaning is given by the back ends.
"""


class StencilModelNode(object):
    _fields = ()

    def __init__(self, **kwargs):
        for name in self._fields:
            setattr(self, name, kwargs.get(name))

    def __repr__(self):
        inner = ", ".join("{}={!r}".format(f, getattr(self, f)) for f in self._fields)
        return "{}({})".format(type(self).__name__, inner)


# ---- the four semantic nodes of the reference -------------------------------------------------
class InteriorPointsLoop(StencilModelNode):
    """``for target in self.interior_points(grid)`` (ref ``stencil_model.py:26-33``)."""
    _fields = ('target', 'body')

    def __init__(self, target=None, body=None):
        super(InteriorPointsLoop, self).__init__(target=target, body=body)


class NeighborPointsLoop(StencilModelNode):
    """``for neighbor_target in self.neighbors(p, neighbor_id)`` (ref ``stencil_model.py:35-45``)."""
    _fields = ('neighbor_id', 'grid_name', 'neighbor_target', 'body')

    def __init__(self, neighbor_id=None, grid_name=None, neighbor_target=None, body=None):
        super(NeighborPointsLoop, self).__init__(
            neighbor_id=neighbor_id, grid_name=grid_name,
            neighbor_target=neighbor_target, body=body)


class MathFunction(StencilModelNode):
    """``self.distance(a, b)`` or ``int(e)`` (ref ``stencil_model.py:48-54``)."""
    _fields = ('func', 'args')

    def __init__(self, func=None, args=None):
        super(MathFunction, self).__init__(func=func, args=args)


class GridElement(StencilModelNode):
    """``grid_name[target]`` (ref ``stencil_model.py:57-75``)."""
    _fields = ('grid_name', 'target')

    def __init__(self, grid_name=None, target=None):
        super(GridElement, self).__init__(grid_name=grid_name, target=target)

    @property
    def name(self):
        return self.grid_name

    @name.setter
    def name(self, value):
        self.grid_name = value

    def copy(self):
        return GridElement(grid_name=self.grid_name, target=self.target)


# ---- plain nodes (ctree's job in the reference) ----------------------------------------------
class KernelFunction(StencilModelNode):
    """The whole ``kernel`` method: parameter names (without self) and statement list."""
    _fields = ('params', 'body')


class Constant(StencilModelNode):
    _fields = ('value',)

    def __init__(self, value=None):
        super(Constant, self).__init__(value=value)


class SymbolRef(StencilModelNode):
    _fields = ('name',)

    def __init__(self, name=None):
        super(SymbolRef, self).__init__(name=name)


class SelfAttribute(StencilModelNode):
    """``self.<name>`` used as a value; resolved against the kernel object at specialisation."""
    _fields = ('name',)

    def __init__(self, name=None):
        super(SelfAttribute, self).__init__(name=name)


class BinaryOp(StencilModelNode):
    _fields = ('op', 'left', 'right')     # op is one of + - * /


class UnaryOp(StencilModelNode):
    _fields = ('op', 'operand')           # op is one of - +


class FunctionCall(StencilModelNode):
    _fields = ('func', 'args')            # func is a plain name, e.g. 'abs', 'min'


class Assign(StencilModelNode):
    _fields = ('target', 'value')


class AugAssign(StencilModelNode):
    _fields = ('target', 'op', 'value')   # op is one of + - * /
