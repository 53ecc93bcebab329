"""Drop-in for the reference's ``timelog`` (timelog.py:3-10): '<ctime> | HH:MM:SS | message'."""
from time import strftime, time


def timelog(message, start_time):
  elapsed = int(time() - start_time)
  clock = '{:02}:{:02}:{:02}'.format(elapsed // 3600, (elapsed % 3600) // 60, elapsed % 60)
  return f"{strftime('%c')} | {clock} | {message}"
