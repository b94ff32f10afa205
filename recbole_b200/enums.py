"""``ModelType`` / ``InputType`` as the reference defines them (recbole/utils/enum_type.py:13-35,73-83).

The reference writes ``model_class.type`` into ``config['MODEL_TYPE']`` (configurator.py:237) and
``model_class.input_type`` into ``config['MODEL_INPUT_TYPE']`` (configurator.py:275-276), then compares them
with its own enum members by ``==``, set membership and dict lookup (configurator.py:238-258,
data/utils.py:40-52,262-286, general_dataloader.py via ``dl_format``).  The fused model classes therefore
carry members that are *equal to* and *hash like* the reference's:

* when ``recbole`` is importable the reference's own classes are used;
* otherwise (the GPU box, the tests) a stand-in with the same names and values is defined whose members compare
  equal to any enum member of a class with the same name and the same value -- so a process that imports
  ``recbole`` later still sees ``FusedBPR.type == ModelType.GENERAL`` and finds it in
  ``{ModelType.GENERAL: 'General', ...}`` (``Enum.__hash__`` is ``hash(name)``, kept here).
"""
from enum import Enum


class _CompatEnum(Enum):
    def __eq__(self, other):
        if self is other:
            return True
        return (isinstance(other, Enum) and type(other).__name__ == type(self).__name__
                and other.name == self.name and other.value == self.value)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._name_)


class _ModelType(_CompatEnum):
    GENERAL = 1
    SEQUENTIAL = 2
    CONTEXT = 3
    KNOWLEDGE = 4
    SOCIAL = 5
    TRADITIONAL = 6
    DECISIONTREE = 7


class _InputType(_CompatEnum):
    POINTWISE = 1
    PAIRWISE = 2
    LISTWISE = 3


_ModelType.__name__ = _ModelType.__qualname__ = "ModelType"
_InputType.__name__ = _InputType.__qualname__ = "InputType"

try:  # the reference's own classes when it is installed / on sys.path
    from recbole.utils.enum_type import InputType, ModelType  # noqa: F401
    FROM_REFERENCE = True
except Exception:  # not importable (or its optional imports are missing): the stand-ins
    ModelType, InputType = _ModelType, _InputType
    FROM_REFERENCE = False
