"""``TypedInput`` / ``BaseWrapper`` -- the plugin API shape the pipeline discovers.

Mirrors /root/reference/wrappers/base_wrapper.py:26-135 (the option descriptor, the singleton and the
abstract ``process_audio``).  The gradio renderers and the FastAPI JSON handler of the reference
(:248-743) are UI / server code and are out of scope.
"""
from __future__ import annotations

import re
from abc import abstractmethod
from typing import Any, Callable, Dict, List, Union

from ..project_files import ProjectFiles


class TypedInput:
    def __init__(self, default: Any = ..., description: str = None, ge: float = None, le: float = None,
                 step: float = None, min_length: int = None, max_length: int = None, regex: str = None,
                 choices: List[Union[str, int]] = None, type: type = None, gradio_type: str = None,
                 render: bool = True, required: bool = False, refresh: Callable = None,
                 on_change: Callable = None, on_click: Callable = None, on_select: Callable = None,
                 controls: List[str] = None, group_name: str = None):
        self.default, self.description, self.ge, self.le, self.step = default, description, ge, le, step
        self.min_length, self.max_length, self.regex, self.choices = min_length, max_length, regex, choices
        self.type, self.render, self.required, self.refresh = type, render, required, refresh
        self.on_change, self.on_click, self.on_select = on_change, on_click, on_select
        self.controls, self.group_name = controls, group_name
        self.gradio_type = gradio_type or self.pick_gradio_type()

    def pick_gradio_type(self) -> str:
        if self.type is bool:
            return "Checkbox"
        if self.type in (int, float) and self.ge is not None and self.le is not None:
            return "Slider"
        if self.type is float:
            return "Number"
        if self.choices:
            return "Dropdown"
        return "Text"

    def validate(self, value: Any) -> Any:
        if self.choices and value not in self.choices:
            raise ValueError(f"{value!r} not in {self.choices}")
        if self.ge is not None and value < self.ge:
            raise ValueError(f"{value!r} < {self.ge}")
        if self.le is not None and value > self.le:
            raise ValueError(f"{value!r} > {self.le}")
        return value


class BaseWrapper:
    _instance = None
    priority = 1000
    allowed_kwargs: Dict[str, TypedInput] = {}
    description = "Base Wrapper"
    default = False
    required = False
    hidden_groups: List[str] = []

    def __new__(cls):
        if cls.__dict__.get("_instance") is None:
            inst = super().__new__(cls)
            if "title" not in cls.__dict__:
                inst.title = " ".join(w.capitalize() for w in re.sub(r"(?<!^)(?=[A-Z])", "_", cls.__name__).split("_"))
            cls._instance = inst
        return cls._instance

    def validate_args(self, **kwargs: Dict[str, Any]) -> bool:
        for key, value in kwargs.items():
            if key in self.allowed_kwargs:
                try:
                    self.allowed_kwargs[key].validate(value)
                except (ValueError, TypeError):
                    return False
        return True

    @abstractmethod
    def process_audio(self, inputs: List[ProjectFiles], callback=None, **kwargs: Dict[str, Any]) -> List[ProjectFiles]:
        raise NotImplementedError
