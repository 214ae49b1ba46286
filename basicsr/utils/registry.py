"""Name -> class registries (same behaviour as the reference's basicsr/utils/registry.py:4-92:
decorator or call registration, duplicate names rejected, ``get`` retries with a ``_basicsr``
suffix)."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj, suffix=None):
        if isinstance(suffix, str):
            name = f"{name}_{suffix}"
        if name in self._obj_map:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None, suffix=None):
        if obj is not None:  # plain call: REGISTRY.register(cls)
            self._do_register(obj.__name__, obj, suffix)
            return None

        def decorator(target):
            self._do_register(target.__name__, target, suffix)
            return target

        return decorator

    def get(self, name, suffix="basicsr"):
        found = self._obj_map.get(name)
        if found is None:
            found = self._obj_map.get(f"{name}_{suffix}")
            if found is not None:
                print(f"Name {name} is not found, use name: {name}_{suffix}!")
        if found is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return found

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())

    def keys(self):
        return self._obj_map.keys()


DATASET_REGISTRY = Registry("dataset")
ARCH_REGISTRY = Registry("arch")
MODEL_REGISTRY = Registry("model")
LOSS_REGISTRY = Registry("loss")
METRIC_REGISTRY = Registry("metric")
