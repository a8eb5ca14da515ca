"""simpy-4.1.1-compatible discrete-event kernel -- ORACLE / TEST INFRASTRUCTURE ONLY.

This is a CPU restatement of the published algorithm of the third-party
dependency that holds the hot path's arithmetic in the reference:
``simpy==4.1.1`` (pinned at reference ``poetry.lock:1234-1235``, declared at
``pyproject.toml:46``).  The upstream sources are NOT vendored in
``/root/reference`` and the wheel is not installable here (no network), so the
behaviour is restated from simpy's documented semantics and
pinned by running the reference's own deterministic hot-path unit tests
(``tests/unit/runtime/**``, ``tests/unit/samplers/**``) unmodified against it
(see ``oracle/run_reference_tests.sh``).

Only the surface AsyncFlow and its tests touch is provided:
``Environment.{now,active_process,schedule,peek,step,run,process,timeout,event}``,
``Event``, ``Timeout``, ``Process``, ``Interrupt``, ``Store``, ``Container``,
``Resource``.

Nothing under ``asyncflow_b200/`` may import this module: the product path is
the CUDA engine.  Importers: ``tests/``, ``oracle/``, ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs and ``__graft_entry__.smoke()``.

Ordering contract (the thing the GPU engine's tie rule is derived from):
the queue holds ``(time, priority, eid, event)`` tuples; ``eid`` is a strictly
increasing insertion counter, so equal ``(time, priority)`` pop FIFO.
``URGENT`` (0) sorts before ``NORMAL`` (1) at equal time.
"""

from __future__ import annotations

from heapq import heappop, heappush
from itertools import count
from typing import Any, Callable, Generator, Iterable, Optional

__version__ = "4.1.1+b200oracle"

__all__ = [
    "Environment", "Event", "Timeout", "Process", "Interrupt", "Store",
    "Container", "Resource", "AllOf", "AnyOf", "URGENT", "NORMAL",
    "EmptySchedule", "StopSimulation", "SimPyException",
]

Infinity = float("inf")
URGENT = 0
NORMAL = 1


class _Pending:
    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return "<PENDING>"


PENDING = _Pending()


class SimPyException(Exception):
    """Base class of the kernel's own exceptions."""


class EmptySchedule(SimPyException):
    """Raised by ``Environment.step`` when no event is left."""


class StopSimulation(SimPyException):
    """Internal: unwinds ``Environment.run`` when the *until* event fires."""

    @classmethod
    def callback(cls, event: "Event") -> None:
        if event._ok:
            raise cls(event._value)
        raise event._value


class Interrupt(SimPyException):
    """Thrown into a process by ``Process.interrupt``."""

    @property
    def cause(self) -> Any:
        return self.args[0]


# --------------------------------------------------------------------------- #
# events                                                                      #
# --------------------------------------------------------------------------- #
class Event:
    """Something that may happen at some point in simulated time.

    *triggered*  <=> a value has been set (the event sits in the queue);
    *processed*  <=> its callbacks have run (``callbacks is None``).
    The two moments are distinct and other events of the same timestamp may be
    processed in between -- ``ServerRuntime`` relies on that
    (reference ``runtime/actors/server.py:212-220``).
    """

    def __init__(self, env: "Environment") -> None:
        self.env = env
        self.callbacks: Optional[list] = []
        self._value: Any = PENDING
        self._ok = True
        self._defused = False

    @property
    def triggered(self) -> bool:
        return self._value is not PENDING

    @property
    def processed(self) -> bool:
        return self.callbacks is None

    @property
    def ok(self) -> bool:
        return self._ok

    @property
    def defused(self) -> bool:
        return self._defused

    @defused.setter
    def defused(self, value: bool) -> None:
        self._defused = bool(value)

    @property
    def value(self) -> Any:
        if self._value is PENDING:
            raise AttributeError(f"Value of {self} is not yet available")
        return self._value

    def trigger(self, event: "Event") -> None:
        self._ok = event._ok
        self._value = event._value
        self.env.schedule(self)

    def succeed(self, value: Any = None) -> "Event":
        if self._value is not PENDING:
            raise RuntimeError(f"{self} has already been triggered")
        self._ok = True
        self._value = value
        self.env.schedule(self)
        return self

    def fail(self, exception: BaseException) -> "Event":
        if self._value is not PENDING:
            raise RuntimeError(f"{self} has already been triggered")
        if not isinstance(exception, BaseException):
            raise TypeError(f"{exception} is not an exception.")
        self._ok = False
        self._value = exception
        self.env.schedule(self)
        return self

    def __and__(self, other: "Event") -> "Condition":
        return Condition(self.env, Condition.all_events, [self, other])

    def __or__(self, other: "Event") -> "Condition":
        return Condition(self.env, Condition.any_events, [self, other])


class Timeout(Event):
    """Triggered at construction, processed ``delay`` later (NORMAL priority)."""

    def __init__(self, env: "Environment", delay: float, value: Any = None) -> None:
        if delay < 0:
            raise ValueError(f"Negative delay {delay}")
        self.env = env
        self.callbacks = []
        self._value = value
        self._delay = delay
        self._ok = True
        self._defused = False
        env.schedule(self, NORMAL, delay)


class Initialize(Event):
    """Starts a process: URGENT at the current time."""

    def __init__(self, env: "Environment", process: "Process") -> None:
        self.env = env
        self.callbacks = [process._resume]
        self._value = None
        self._ok = True
        self._defused = False
        env.schedule(self, URGENT)


class Interruption(Event):
    def __init__(self, process: "Process", cause: Any) -> None:
        self.env = process.env
        self.callbacks = [self._interrupt]
        self._value = Interrupt(cause)
        self._ok = False
        self._defused = True
        if process._value is not PENDING:
            raise RuntimeError(f"{process} has terminated and cannot be interrupted.")
        if process is self.env.active_process:
            raise RuntimeError("A process is not allowed to interrupt itself.")
        self.process = process
        self.env.schedule(self, URGENT)

    def _interrupt(self, event: Event) -> None:
        if self.process._value is not PENDING:
            return
        self.process._target.callbacks.remove(self.process._resume)
        self.process._resume(self)


class Process(Event):
    """A generator driven by the events it yields; itself an event (its exit)."""

    def __init__(self, env: "Environment", generator: Generator) -> None:
        if not hasattr(generator, "throw"):
            raise ValueError(f"{generator} is not a generator.")
        self.env = env
        self.callbacks = []
        self._value = PENDING
        self._ok = True
        self._defused = False
        self._generator = generator
        self._target: Event = Initialize(env, self)

    @property
    def target(self) -> Event:
        return self._target

    @property
    def name(self) -> str:
        return self._generator.__name__

    @property
    def is_alive(self) -> bool:
        return self._value is PENDING

    def interrupt(self, cause: Any = None) -> None:
        Interruption(self, cause)

    def _resume(self, event: Event) -> None:
        env = self.env
        env._active_proc = self
        while True:
            try:
                if event._ok:
                    nxt = self._generator.send(event._value)
                else:
                    event._defused = True
                    exc = type(event._value)(*event._value.args)
                    exc.__cause__ = event._value
                    nxt = self._generator.throw(exc)
            except StopIteration as stop:
                nxt = None
                self._ok = True
                self._value = stop.args[0] if len(stop.args) else None
                env.schedule(self)
                break
            except BaseException as exc:  # noqa: BLE001 - mirrors simpy
                nxt = None
                self._ok = False
                self._value = exc
                env.schedule(self)
                break
            try:
                if nxt.callbacks is not None:
                    # not processed yet: park until it is
                    nxt.callbacks.append(self._resume)
                    break
            except AttributeError:
                msg = f'Invalid yield value "{nxt}"'
                err = RuntimeError(msg)
                err.__cause__ = None
                # feed the error back into the generator on the next spin
                event = Event(env)
                event._ok = False
                event._value = err
                continue
            # already processed: feed its value straight back in
            event = nxt
        self._target = nxt
        env._active_proc = None


class ConditionValue:
    def __init__(self) -> None:
        self.events: list[Event] = []

    def __getitem__(self, key: Event) -> Any:
        if key not in self.events:
            raise KeyError(str(key))
        return key._value

    def __contains__(self, key: Event) -> bool:
        return key in self.events

    def __eq__(self, other: object) -> bool:
        if isinstance(other, ConditionValue):
            return self.events == other.events
        return self.todict() == other

    def __iter__(self):
        return iter(self.events)

    def keys(self):
        return iter(self.events)

    def values(self):
        return (e._value for e in self.events)

    def items(self):
        return ((e, e._value) for e in self.events)

    def todict(self) -> dict:
        return {e: e._value for e in self.events}


class Condition(Event):
    def __init__(self, env: "Environment", evaluate: Callable, events: Iterable[Event]) -> None:
        super().__init__(env)
        self._evaluate = evaluate
        self._events = tuple(events)
        self._count = 0
        if not self._events:
            self.succeed(ConditionValue())
            return
        for ev in self._events:
            if ev.env is not env:
                raise ValueError("It is not allowed to mix events from different environments")
        for ev in self._events:
            if ev.callbacks is None:
                self._check(ev)
            else:
                ev.callbacks.append(self._check)
        assert isinstance(self.callbacks, list)
        self.callbacks.append(self._build_value)

    def _populate_value(self, value: ConditionValue) -> None:
        for ev in self._events:
            if isinstance(ev, Condition):
                ev._populate_value(value)
            elif ev.callbacks is None:
                value.events.append(ev)

    def _build_value(self, event: Event) -> None:
        self._remove_check_callbacks()
        if event._ok:
            self._value = ConditionValue()
            self._populate_value(self._value)

    def _remove_check_callbacks(self) -> None:
        for ev in self._events:
            if ev.callbacks and self._check in ev.callbacks:
                ev.callbacks.remove(self._check)
            if isinstance(ev, Condition):
                ev._remove_check_callbacks()

    def _check(self, event: Event) -> None:
        if self._value is not PENDING:
            return
        self._count += 1
        if not event._ok:
            event._defused = True
            self.fail(event._value)
        elif self._evaluate(self._events, self._count):
            self.succeed()

    @staticmethod
    def all_events(events: tuple, count: int) -> bool:
        return len(events) == count

    @staticmethod
    def any_events(events: tuple, count: int) -> bool:
        return count > 0 or len(events) == 0


class AllOf(Condition):
    def __init__(self, env: "Environment", events: Iterable[Event]) -> None:
        super().__init__(env, Condition.all_events, events)


class AnyOf(Condition):
    def __init__(self, env: "Environment", events: Iterable[Event]) -> None:
        super().__init__(env, Condition.any_events, events)


# --------------------------------------------------------------------------- #
# environment                                                                 #
# --------------------------------------------------------------------------- #
class Environment:
    """Next-event time advance over a binary heap of ``(t, prio, eid, event)``."""

    def __init__(self, initial_time: float = 0) -> None:
        self._now = initial_time
        self._queue: list = []
        self._eid = count()
        self._active_proc: Optional[Process] = None

    @property
    def now(self) -> float:
        return self._now

    @property
    def active_process(self) -> Optional[Process]:
        return self._active_proc

    # factories ------------------------------------------------------------- #
    def process(self, generator: Generator) -> Process:
        return Process(self, generator)

    def timeout(self, delay: float = 0, value: Any = None) -> Timeout:
        return Timeout(self, delay, value)

    def event(self) -> Event:
        return Event(self)

    def all_of(self, events: Iterable[Event]) -> AllOf:
        return AllOf(self, events)

    def any_of(self, events: Iterable[Event]) -> AnyOf:
        return AnyOf(self, events)

    # queue ----------------------------------------------------------------- #
    def schedule(self, event: Event, priority: int = NORMAL, delay: float = 0) -> None:
        heappush(self._queue, (self._now + delay, priority, next(self._eid), event))

    def peek(self) -> float:
        try:
            return self._queue[0][0]
        except IndexError:
            return Infinity

    def step(self) -> None:
        try:
            self._now, _, _, event = heappop(self._queue)
        except IndexError:
            raise EmptySchedule from None
        callbacks, event.callbacks = event.callbacks, None
        for cb in callbacks:
            cb(event)
        if not event._ok and not event._defused:
            exc = type(event._value)(*event._value.args)
            exc.__cause__ = event._value
            raise exc

    def run(self, until: Any = None) -> Any:
        if until is not None:
            if not isinstance(until, Event):
                at = float(until)
                if at <= self._now:
                    raise ValueError(f"until ({at}) must be greater than the current simulation time")
                until = Event(self)
                until._ok = True
                until._value = None
                # URGENT: NORMAL events at exactly `at` are NOT processed
                self.schedule(until, URGENT, at - self._now)
            elif until.callbacks is None:
                return until.value
            until.callbacks.append(StopSimulation.callback)
        try:
            while True:
                self.step()
        except StopSimulation as stop:
            return stop.args[0]
        except EmptySchedule:
            if until is not None:
                assert not until.triggered
                raise RuntimeError(
                    f'No scheduled events left but "until" event was not triggered: {until}'
                ) from None
        return None


# --------------------------------------------------------------------------- #
# shared resources                                                            #
# --------------------------------------------------------------------------- #
class Put(Event):
    def __init__(self, resource: "BaseResource") -> None:
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.put_queue.append(self)
        self.callbacks.append(resource._trigger_get)
        resource._trigger_put(None)

    def __enter__(self) -> "Put":
        return self

    def __exit__(self, *exc: Any) -> Optional[bool]:
        self.cancel()
        return None

    def cancel(self) -> None:
        if not self.triggered:
            self.resource.put_queue.remove(self)


class Get(Event):
    def __init__(self, resource: "BaseResource") -> None:
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.get_queue.append(self)
        self.callbacks.append(resource._trigger_put)
        resource._trigger_get(None)

    def __enter__(self) -> "Get":
        return self

    def __exit__(self, *exc: Any) -> Optional[bool]:
        self.cancel()
        return None

    def cancel(self) -> None:
        if not self.triggered:
            self.resource.get_queue.remove(self)


class BaseResource:
    """put/get queues re-examined FIFO; a falsy ``_do_*`` stops the walk."""

    def __init__(self, env: Environment, capacity: float) -> None:
        self._env = env
        self._capacity = capacity
        self.put_queue: list = []
        self.get_queue: list = []

    @property
    def capacity(self) -> float:
        return self._capacity

    def _do_put(self, event: Put) -> Optional[bool]:
        raise NotImplementedError

    def _do_get(self, event: Get) -> Optional[bool]:
        raise NotImplementedError

    def _trigger_put(self, get_event: Optional[Get]) -> None:
        idx = 0
        while idx < len(self.put_queue):
            put_event = self.put_queue[idx]
            proceed = self._do_put(put_event)
            if not put_event.triggered:
                idx += 1
            elif self.put_queue.pop(idx) != put_event:
                raise RuntimeError("Put queue invariant violated")
            if not proceed:
                break

    def _trigger_get(self, put_event: Optional[Put]) -> None:
        idx = 0
        while idx < len(self.get_queue):
            get_event = self.get_queue[idx]
            proceed = self._do_get(get_event)
            if not get_event.triggered:
                idx += 1
            elif self.get_queue.pop(idx) != get_event:
                raise RuntimeError("Get queue invariant violated")
            if not proceed:
                break


class ContainerPut(Put):
    def __init__(self, container: "Container", amount: float) -> None:
        if amount <= 0:
            raise ValueError(f"amount(={amount}) must be > 0.")
        self.amount = amount
        super().__init__(container)


class ContainerGet(Get):
    def __init__(self, container: "Container", amount: float) -> None:
        if amount <= 0:
            raise ValueError(f"amount(={amount}) must be > 0.")
        self.amount = amount
        super().__init__(container)


class Container(BaseResource):
    """Counting resource: strict FIFO with head-of-line blocking."""

    def __init__(self, env: Environment, capacity: float = Infinity, init: float = 0) -> None:
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        if init < 0:
            raise ValueError('"init" must be >= 0.')
        if init > capacity:
            raise ValueError('"init" must be <= "capacity".')
        super().__init__(env, capacity)
        self._level = init

    @property
    def level(self) -> float:
        return self._level

    def put(self, amount: float) -> ContainerPut:
        return ContainerPut(self, amount)

    def get(self, amount: float) -> ContainerGet:
        return ContainerGet(self, amount)

    def _do_put(self, event: ContainerPut) -> Optional[bool]:
        if self._capacity - self._level >= event.amount:
            self._level += event.amount
            event.succeed()
            return True
        return None

    def _do_get(self, event: ContainerGet) -> Optional[bool]:
        if self._level >= event.amount:
            self._level -= event.amount
            event.succeed()
            return True
        return None


class StorePut(Put):
    def __init__(self, store: "Store", item: Any) -> None:
        self.item = item
        super().__init__(store)


class StoreGet(Get):
    pass


class Store(BaseResource):
    """FIFO mailbox; each trigger pass serves at most one request."""

    def __init__(self, env: Environment, capacity: float = Infinity) -> None:
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        super().__init__(env, capacity)
        self.items: list = []

    def put(self, item: Any) -> StorePut:
        return StorePut(self, item)

    def get(self) -> StoreGet:
        return StoreGet(self)

    def _do_put(self, event: StorePut) -> Optional[bool]:
        if len(self.items) < self._capacity:
            self.items.append(event.item)
            event.succeed()
        return None

    def _do_get(self, event: StoreGet) -> Optional[bool]:
        if self.items:
            event.succeed(self.items.pop(0))
        return None


class _Request(Put):
    def __init__(self, resource: "Resource") -> None:
        self.usage_since: Optional[float] = None
        super().__init__(resource)

    def __exit__(self, *exc: Any) -> Optional[bool]:
        super().__exit__(*exc)
        if exc[0] is not GeneratorExit:
            self.resource.release(self)
        return None


class _Release(Get):
    def __init__(self, resource: "Resource", request: _Request) -> None:
        self.request = request
        super().__init__(resource)


class Resource(BaseResource):
    """Classic ``capacity``-slot resource (not used by AsyncFlow; kept for tests)."""

    def __init__(self, env: Environment, capacity: int = 1) -> None:
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        super().__init__(env, capacity)
        self.users: list = []
        self.queue = self.put_queue

    @property
    def count(self) -> int:
        return len(self.users)

    def request(self) -> _Request:
        return _Request(self)

    def release(self, request: _Request) -> _Release:
        return _Release(self, request)

    def _do_put(self, event: _Request) -> Optional[bool]:
        if len(self.users) < self._capacity:
            self.users.append(event)
            event.usage_since = self._env.now
            event.succeed()
        return None

    def _do_get(self, event: _Release) -> Optional[bool]:
        try:
            self.users.remove(event.request)
        except ValueError:
            pass
        event.succeed()
        return None
