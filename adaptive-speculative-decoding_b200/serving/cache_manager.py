"""``KVCacheManager``: the reference's per-(request, stage) result store with byte accounting, LRU
eviction and a TTL sweep (/root/reference/src/serving/cache_manager.py:26-387).  Despite its name it
never held K/V tensors; the paged K/V pool of the B200 engine lives in ``asd_b200.engine.QwenEngine``.
Same public API; unlike the reference it tolerates the non-tensor payloads the pipeline stores
(the reference calls ``.is_cuda`` on ``str`` / ``ndarray`` and crashes, cache_manager.py:176-178,213-215)."""
from __future__ import annotations

import logging
import threading
import time
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Dict, Optional

logger = logging.getLogger(__name__)


@dataclass
class CacheEntry:
    cache_data: Dict[str, Any]
    creation_time: float
    last_access: float
    size_bytes: int
    stage_id: int


def _nbytes(v) -> int:
    if hasattr(v, "element_size") and hasattr(v, "numel"):
        return int(v.element_size() * v.numel())
    if hasattr(v, "nbytes"):
        return int(v.nbytes)
    if isinstance(v, (str, bytes)):
        return len(v)
    return 0


class KVCacheManager:
    def __init__(self, num_stages: int = 4, max_cache_size_gb: float = 40.0, cleanup_interval: float = 300.0,
                 enable_compression: bool = False, max_age_seconds: float = 1800.0):
        self.num_stages = num_stages
        self.max_cache_size_bytes = int(max_cache_size_gb * (1024 ** 3))
        self.cleanup_interval = cleanup_interval
        self.enable_compression = enable_compression
        self.max_age_seconds = max_age_seconds
        self.caches: Dict[str, Dict[int, CacheEntry]] = defaultdict(dict)
        self.stats = {"total_allocations": 0, "total_deallocations": 0, "current_size_bytes": 0, "cache_hits": 0,
                      "cache_misses": 0, "cleanup_count": 0}
        self.lock = threading.RLock()
        self._stop = threading.Event()
        self.cleanup_thread = threading.Thread(target=self._periodic_cleanup, daemon=True)
        self.cleanup_thread.start()

    # -- public API (cache_manager.py:69,121,149,192,369)
    def allocate(self, request_id: str, stage_id: int, cache_data: Dict[str, Any]) -> bool:
        with self.lock:
            size = sum(_nbytes(v) for v in cache_data.values())
            if not self._ensure_space(size):
                logger.warning("Insufficient cache space for request %s", request_id)
                return False
            old = self.caches[request_id].get(stage_id)
            if old is not None:
                self.stats["current_size_bytes"] -= old.size_bytes
            now = time.time()
            self.caches[request_id][stage_id] = CacheEntry(cache_data, now, now, size, stage_id)
            self.stats["total_allocations"] += 1
            self.stats["current_size_bytes"] += size
            return True

    def get_cache(self, request_id: str, stage_id: int) -> Optional[Dict[str, Any]]:
        with self.lock:
            entry = self.caches.get(request_id, {}).get(stage_id)
            if entry is None:
                self.stats["cache_misses"] += 1
                return None
            entry.last_access = time.time()
            self.stats["cache_hits"] += 1
            return entry.cache_data

    def truncate_at_stage(self, request_id: str, stage_id: int):
        """drop the entries of stages after ``stage_id`` (the cascade stopped there)"""
        with self.lock:
            stages = self.caches.get(request_id)
            if not stages:
                return
            for s in [s for s in stages if s > stage_id]:
                self._drop(request_id, s)

    def cleanup_request(self, request_id: str):
        with self.lock:
            for s in list(self.caches.get(request_id, {})):
                self._drop(request_id, s)
            self.caches.pop(request_id, None)

    def get_stats(self) -> Dict[str, Any]:
        with self.lock:
            hits, miss = self.stats["cache_hits"], self.stats["cache_misses"]
            return {**self.stats, "active_requests": len(self.caches),
                    "active_stages": sum(len(s) for s in self.caches.values()),
                    "current_size_mb": self.stats["current_size_bytes"] / (1024 ** 2),
                    "max_size_mb": self.max_cache_size_bytes / (1024 ** 2),
                    "utilization": self.stats["current_size_bytes"] / self.max_cache_size_bytes,
                    "hit_rate": hits / (hits + miss) if hits + miss > 0 else 0}

    def shutdown(self):
        self._stop.set()

    # -- internals
    def _drop(self, request_id: str, stage_id: int):
        entry = self.caches[request_id].pop(stage_id, None)
        if entry is not None:
            self.stats["current_size_bytes"] -= entry.size_bytes
            self.stats["total_deallocations"] += 1

    def _ensure_space(self, need: int) -> bool:
        if need > self.max_cache_size_bytes:
            return False
        while self.stats["current_size_bytes"] + need > self.max_cache_size_bytes:
            if not self._evict_lru():
                return False
        return True

    def _evict_lru(self) -> bool:
        oldest = None
        for rid, stages in self.caches.items():
            for sid, e in stages.items():
                if oldest is None or e.last_access < oldest[2]:
                    oldest = (rid, sid, e.last_access)
        if oldest is None:
            return False
        self._drop(oldest[0], oldest[1])
        if not self.caches[oldest[0]]:
            self.caches.pop(oldest[0], None)
        return True

    def _periodic_cleanup(self):
        while not self._stop.wait(self.cleanup_interval):
            self._cleanup_expired()

    def _cleanup_expired(self, max_age_seconds: Optional[float] = None):
        limit = time.time() - (self.max_age_seconds if max_age_seconds is None else max_age_seconds)
        with self.lock:
            for rid in list(self.caches):
                for sid in [s for s, e in self.caches[rid].items() if e.last_access < limit]:
                    self._drop(rid, sid)
                if not self.caches[rid]:
                    self.caches.pop(rid, None)
            self.stats["cleanup_count"] += 1
