"""Status-log mixin with the reference's surface (tools/loggingclass.py:10-22): a class-level list shared by
every subclass, drained by ``get_new_status_logs``."""


class GlobalLoggingClass:
    status_logs = []

    def get_new_status_logs(self):
        logs = list(self.status_logs)
        self.status_logs.clear()
        return logs
