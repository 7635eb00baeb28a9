"""CPU: RenderPackage behaves like the plain dict the reference's render() returns
(gaussian_renderer/__init__.py:147-155) while deferring the two lazy entries."""
from curve_gaussian_b200.renderer import RenderPackage


def test_lazy_entries_look_present_and_run_once():
    calls = []

    def mk(name, val):
        def f():
            calls.append(name)
            return val
        return f

    p = RenderPackage({"render": 1, "radii": 2}, {"visibility_filter": mk("vis", 3), "rend_dir": mk("dir", 4)})
    assert "visibility_filter" in p and "rend_dir" in p and "nope" not in p
    assert len(p) == 4 and calls == []
    assert p["render"] == 1 and calls == []
    assert p["visibility_filter"] == 3 and p["visibility_filter"] == 3 and calls == ["vis"]
    assert p.get("rend_dir") == 4 and p.get("nope", 9) == 9 and calls == ["vis", "dir"]
    assert set(p.keys()) == {"render", "radii", "visibility_filter", "rend_dir"}


def test_iteration_forces_everything():
    p = RenderPackage({"a": 1}, {"b": lambda: 2})
    assert dict(p.items()) == {"a": 1, "b": 2}
    q = RenderPackage({"a": 1}, {"b": lambda: 2})
    assert sorted(q) == ["a", "b"] and sorted(q.values()) == [1, 2]
    try:
        RenderPackage({}, {})["missing"]
        raise AssertionError("KeyError expected")
    except KeyError:
        pass
