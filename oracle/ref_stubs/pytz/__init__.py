import datetime

utc = datetime.timezone.utc
